// Shared by the fused blend+skinning kernels (k_body_tc.cu, k_body_pair.cu): tile constants and
// the verts store routine.
#pragma once
#ifndef FB_ABLATE
#define FB_ABLATE 0
#endif
// evict-first stores: verts are written once and never re-read here (.cg, .wt and plain stores: +7 %)
#define FB_ST(p, v) __stcs((p), (v))
#define FB_VT 128                     // vertices per super-tile (MMA M)
#define FB_D_BYTES (FB_VT * 128)      // one Dt16 k-block of one plane: 16 KB
#define FB_W_BYTES (FB_VT * 128)      // W16 tile: 16 KB


// The 12 stores of a warp's 4 samples (verts[b][v][xyz]: three scalar stores 12 B apart per sample; rows
// are only 8-byte aligned, so nothing wider) as a real call.  Inlined into the epilogue, the compiler
// runs out of uniform registers, keeps the 64-bit global-memory descriptor that every STG needs in
// a vector register pair and converts it back with two R2UR per store: 15 % of all instructions
// executed, on the 16-cycle XU pipe, which ncu showed 88 % busy.  In its own frame the descriptor
// is one uniform load.
static __device__ __noinline__ void store_rows4(float *dst, int row_stride, int rows_left, bool v_ok, float a0, float a1, float a2,
                                         float b0, float b1, float b2, float c0, float c1, float c2, float d0, float d1,
                                         float d2) {
  if (!v_ok) return;
#if FB_ABLATE == 3
  return;   // tuning build: no stores at all (everything but the store path)
#endif
#if FB_ABLATE == 5
  // tuning build: same stores and bytes, but a warp's three stores of a row each cover 128 contiguous
  // bytes (WRONG element order) -- what fully coalesced store requests would buy
  {
    const int ln = threadIdx.x & 31;
    dst -= 2 * ln;
    if (rows_left > 0) { FB_ST(dst, a0); FB_ST(dst + 32, a1); FB_ST(dst + 64, a2); }
    if (rows_left > 1) { FB_ST(dst + row_stride, b0); FB_ST(dst + row_stride + 32, b1); FB_ST(dst + row_stride + 64, b2); }
    if (rows_left > 2) { FB_ST(dst + 2 * row_stride, c0); FB_ST(dst + 2 * row_stride + 32, c1); FB_ST(dst + 2 * row_stride + 64, c2); }
    if (rows_left > 3) { FB_ST(dst + 3 * row_stride, d0); FB_ST(dst + 3 * row_stride + 32, d1); FB_ST(dst + 3 * row_stride + 64, d2); }
    return;
  }
#endif
  if (rows_left > 0) {
    FB_ST(dst, a0);
    FB_ST(dst + 1, a1);
    FB_ST(dst + 2, a2);
  }
  if (rows_left > 1) {
    FB_ST(dst + row_stride, b0);
    FB_ST(dst + row_stride + 1, b1);
    FB_ST(dst + row_stride + 2, b2);
  }
  if (rows_left > 2) {
    FB_ST(dst + 2 * row_stride, c0);
    FB_ST(dst + 2 * row_stride + 1, c1);
    FB_ST(dst + 2 * row_stride + 2, c2);
  }
  if (rows_left > 3) {
    FB_ST(dst + 3 * row_stride, d0);
    FB_ST(dst + 3 * row_stride + 1, d1);
    FB_ST(dst + 3 * row_stride + 2, d2);
  }
}

