// Shared by the loss kernels (k_loss.cu) and the lattice nearest-neighbour search (k_mesh_lattice.cu).
#pragma once
#include "smplb_internal.h"

// Fixed-order block sum of one float per thread (blockDim.x a power of two <= 1024).
__device__ __forceinline__ float block_sum(float v, float *red) {
  int t = threadIdx.x;
  red[t] = v;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  float r = red[0];
  __syncthreads();
  return r;
}

// d2(a, b) exactly as the fp32 expansion of ops.py:63-65: (-2 a.b + |a|^2) + |b|^2, each sum
// rounded separately (-2 * x is exact, so the FMA below rounds once like the TF add does).
__device__ __forceinline__ float d2_expand(float ax, float ay, float a2, float bx, float by, float b2) {
  float dot = __fmaf_rn(ax, bx, __fmul_rn(ay, by));
  return __fadd_rn(__fmaf_rn(-2.0f, dot, a2), b2);
}

#define GP_STRIDE 8   // per (image, set) grid parameters: x0, y0, x1, y1, inv_h, h, max |p|^2, -

#define LAT_N 256                       // lattice extent (the reference's images are 224 x 224)
#define LAT_W (LAT_N / 32)              // 32-bit words per row
#define LAT_C 64                        // coarse cells per side (4 x 4 pixels each)
#define LAT_BM_WORDS (LAT_N * LAT_W)
#define LAT_BYTES (2 * LAT_BM_WORDS * 4 + LAT_N * 4 + LAT_C * LAT_C)   // bitmap, transposed bitmap, row prefix, coarse distances



// Chebyshev distance, in cells, from every cell of a 64 x 64 grid to the nearest occupied one (64 if none is):
// out[64 y + x] = number of 3 x 3 dilations of the occupancy mask that it takes to reach cell (x, y).  occ[y] holds row y
// (bit x = cell occupied; destroyed), tmp[64] is scratch; rows are walked by threads 0..63, the whole CTA (256 threads)
// must call.  (The first version took, per cell, the minimum over rows of max(row distance, distance along that row)
// with an early exit: 56 % of the build kernel's time for the cells far from every point.)
__device__ __forceinline__ void chebyshev_cells64(unsigned long long *occ, unsigned long long *tmp, unsigned char *out) {
  const int t = threadIdx.x;
  for (int k = t; k < 64 * 64; k += 256) out[k] = (occ[k >> 6] >> (k & 63) & 1ull) ? 0 : 64;
  // (barrier: the fill above is ordered before the rows' own writes below; nothing occupied: every cell keeps 64)
  if (__syncthreads_or(t < 64 && occ[t] != 0ull) == 0) return;
  unsigned long long *cur = occ, *nxt = tmp;
  for (int k = 1; k < 64; ++k) {
    bool open = false;
    if (t < 64) {
      const unsigned long long a = cur[t];
      const unsigned long long u = a | (t > 0 ? cur[t - 1] : 0ull) | (t < 63 ? cur[t + 1] : 0ull);
      const unsigned long long d = u | (u << 1) | (u >> 1);
      nxt[t] = d;
      for (unsigned long long nw = d & ~a; nw; nw &= nw - 1) out[64 * t + __ffsll((long long)nw) - 1] = (unsigned char)k;
      open = d != ~0ull;
    }
    const int more = __syncthreads_or(open);
    unsigned long long *sw = cur;
    cur = nxt;
    nxt = sw;
    if (!more) break;
  }
  __syncthreads();
}

// The same for the coarse cells (4 x 4 lattice points each) of a row-major 256 x 256 bitmap in shared memory.
__device__ __forceinline__ void coarse_distance(const unsigned *bm, unsigned long long *s_occ, unsigned long long *s_tmp,
                                                unsigned char *out_cd) {
  const int t = threadIdx.x;
  {
    // thread t: 16 cells of row t / 4
    const int y = t >> 2, c0 = (t & 3) * 16;
    unsigned m = 0;
    for (int cx = c0; cx < c0 + 16; ++cx) {
      unsigned any = 0;
      for (int r = 0; r < 4; ++r) any |= (bm[(4 * y + r) * LAT_W + (cx >> 3)] >> (4 * (cx & 7))) & 0xFu;
      m |= (unsigned)(any != 0) << (cx - c0);
    }
    reinterpret_cast<unsigned short *>(s_occ)[4 * y + (t & 3)] = (unsigned short)m;   // little-endian: bits c0 .. c0 + 15 of row y
  }
  __syncthreads();
  chebyshev_cells64(s_occ, s_tmp, out_cd);
}

// Pixel tables of image i for the lattice search (k_mesh_lattice.cu), built by the 256 threads of one CTA: bitmap,
// transposed bitmap, row prefix, coarse distances (LAT_BYTES at lat + i * LAT_BYTES), the bounding box of the pixels
// (gparam of set 1, as k_grid_build writes it) and lat_ok[i] = the points ARE a row-major pixel list (integers in
// [0, LAT_N)^2, strictly increasing).  Returns lat_ok[i] to every thread.
// Scratch (shared memory of the caller): bm, bmT [LAT_BM_WORDS] each, scan [256], s_occ, s_tmp [LAT_C] each.
__device__ __forceinline__ bool lattice_build_image(int i, const float *__restrict__ pts, const int *__restrict__ offsets,
                                                    unsigned char *__restrict__ lat, int *__restrict__ lat_ok,
                                                    float *__restrict__ gparam, unsigned *bm, unsigned *bmT, int *scan,
                                                    unsigned long long *s_occ, unsigned long long *s_tmp) {
  __shared__ int s_bad, s_box[4];
  const int t = threadIdx.x;
  const int p0 = offsets[i], np = offsets[i + 1] - p0;
  for (int k = t; k < LAT_BM_WORDS; k += 256) bm[k] = bmT[k] = 0u;
  if (t == 0) {
    s_bad = 0;
    s_box[0] = s_box[1] = LAT_N;
    s_box[2] = s_box[3] = -1;
  }
  __syncthreads();
  int bad = 0;
  {
    // 8-byte point loads, four iterations in flight; the previous point of the list comes from the neighbouring lane
    const float2 *p2 = reinterpret_cast<const float2 *>(pts) + p0;
    const int lane = t & 31;
#pragma unroll 4
    for (int k0 = 0; k0 < np; k0 += 256) {
      const int k = k0 + t;
      const bool in = k < np;
      const float2 pt = in ? p2[k] : make_float2(0.f, 0.f);
      float px = __shfl_up_sync(0xffffffffu, pt.x, 1), py = __shfl_up_sync(0xffffffffu, pt.y, 1);
      if (lane == 0 && in && k > 0) {
        const float2 pv = p2[k - 1];
        px = pv.x;
        py = pv.y;
      }
      if (!in) continue;
      const float x = pt.x, y = pt.y;
      const int xi = (int)x, yi = (int)y;
      if (!(x == (float)xi && y == (float)yi && xi >= 0 && xi < LAT_N && yi >= 0 && yi < LAT_N)) {
        bad = 1;
        continue;
      }
      // strictly increasing (row, col): the list is the bitmap's own order, without duplicates
      if (k > 0 && !(py < y || (py == y && px < x))) bad = 1;
      atomicOr(&bm[yi * LAT_W + (xi >> 5)], 1u << (xi & 31));
      atomicOr(&bmT[xi * LAT_W + (yi >> 5)], 1u << (yi & 31));
    }
  }
  if (bad) s_bad = 1;
  __syncthreads();
  unsigned char *out = lat + (size_t)i * LAT_BYTES;
  unsigned *o_bm = reinterpret_cast<unsigned *>(out), *o_bmT = o_bm + LAT_BM_WORDS;
  int *o_pref = reinterpret_cast<int *>(o_bmT + LAT_BM_WORDS);
  unsigned char *o_cd = reinterpret_cast<unsigned char *>(o_pref + LAT_N);
  for (int k = t; k < LAT_BM_WORDS; k += 256) {
    o_bm[k] = bm[k];
    o_bmT[k] = bmT[k];
  }
  // row prefix: pixels in the rows above (thread = row); bounding box from the non-empty rows / columns
  int cs = 0, ccol = 0;
  for (int w = 0; w < LAT_W; ++w) {
    cs += __popc(bm[t * LAT_W + w]);
    ccol |= bmT[t * LAT_W + w] != 0u;
  }
  if (cs) {
    atomicMin(&s_box[1], t);
    atomicMax(&s_box[3], t);
  }
  if (ccol) {
    atomicMin(&s_box[0], t);
    atomicMax(&s_box[2], t);
  }
  scan[t] = cs;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    int v = t >= o ? scan[t - o] : 0;
    __syncthreads();
    scan[t] += v;
    __syncthreads();
  }
  o_pref[t] = scan[t] - cs;
  if (t == 255 && scan[255] != np) s_bad = 1;
  coarse_distance(bm, s_occ, s_tmp, o_cd);
  if (t == 0) {
    lat_ok[i] = s_bad ? 0 : 1;
    float *gp = gparam + ((size_t)i * 2 + 1) * GP_STRIDE;
    const bool any = s_box[2] >= 0;
    const float x0 = any ? (float)s_box[0] : 0.f, y0 = any ? (float)s_box[1] : 0.f;
    const float x1 = any ? (float)s_box[2] : 0.f, y1 = any ? (float)s_box[3] : 0.f;
    gp[0] = x0;
    gp[1] = y0;
    gp[2] = x1;
    gp[3] = y1;
    gp[4] = 1.0f;
    gp[5] = 1.0f;
    gp[6] = x1 * x1 + y1 * y1;   // max |a|^2 (coordinates are >= 0)
    gp[7] = 0.f;
  }
  __syncthreads();
  return s_bad == 0;
}

