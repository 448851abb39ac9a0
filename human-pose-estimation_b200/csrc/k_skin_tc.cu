// Linear blend skinning with the per-vertex transforms formed on the tensor cores.
//
//   T[v, (s, e)] = sum_j W[v, j] * A[s, j, e]        e = 4r + d indexes the 3x4 transform
//   verts[s, v, r] = T[v,s,4r+0..2] . v_posed[s, v, :] + T[v,s,4r+3]     (batch_smpl.py:139-149)
//
// Dense 24-wide blending costs 288 FMA per (vertex, sample) on the FP32 pipes -- 2-4x more time
// than moving the kernel's bytes -- so T runs as a tcgen05 GEMM: M = 128 vertices (TMEM lanes),
// N = 12 * 16 samples, K = 72 -> 80.  fp16 operands are split so the product keeps fp32-grade
// accuracy (a single fp16/tf32 rounding of W would move vertices by ~0.5 mm, 5x the tolerance):
//   T = W_hi x A_hi + W_hi x A_lo + W_lo x A_hi.
// Both operands are ONE 64-column (128 B) swizzle atom per row, organised in four 16-column
// windows so that five K = 16 MMAs, each pairing a window of W16 with a window of A16, produce
// exactly those 72 products (16 B-aligned window starts are all the descriptor needs):
//   window      0             1                          2              3
//   W16     W_hi[0:16]   W_hi[16:24] W_hi[16:24]     W_lo[0:16]    W_lo[16:24] 0
//   A16     A_hi[0:16]   A_hi[16:24] A_lo[16:24]     A_lo[0:16]    0
//   MMAs    (W0,A0) (W1,A1) (W0,A2) (W2,A0) (W3,A1)
// Halving the operand rows (they were two atoms) cut the L2 -> SM traffic of a tile from 80 KB
// to 64 KB, which is what this kernel is bound by once HBM keeps up.
// T never leaves the SM: the epilogue threads (one per vertex = TMEM lane) read it with
// tcgen05.ld, apply it to v_posed (staged by TMA) and write verts.  The kernel is bound by those
// two streams.
//
// Persistent, warp-specialised: warp 0 TMA producer of the MMA operands (W16 tile per vertex
// tile, A16 chunk per tile), warp 1 MMA issuer + TMEM allocator, warps 2-9 epilogue, warp 10 TMA
// producer of the v_posed tiles (a 3-deep ring, so ~70 KB of loads are in flight per SM: the
// epilogue never waits on an HBM round trip); two TMEM accumulator stages.
#include <cuda.h>
#include <cuda_fp16.h>

#include "smplb_internal.h"
#include "tc_ptx.cuh"

#define ST_VT 128                 // vertices per tile (MMA M)
#define ST_S 16                   // samples per tile
#define ST_N (12 * ST_S)          // MMA N = 192
#define ST_KP 64                  // operand row: one 64-wide swizzle atom
#define ST_ASTAGES 2
#define ST_PSTAGES 4
#define ST_THREADS 352
#define ST_W_BYTES (ST_VT * 128)          // 16 KB: 128 rows x 128 B
#define ST_A_BYTES (ST_N * 128)           // 24 KB: the A16 chunk of 16 samples
#define ST_SM_A 0                                            // resident A16 chunk (reloaded when the sample chunk changes)
#define ST_SM_W (ST_A_BYTES)                                 // W16 tile ring
#define ST_P_BYTES (3 * ST_S * ST_VT * 4)                    // 24 KB: v_posed tile [xyz][16 samples][128 vertices]
#define ST_SM_P (ST_SM_W + ST_ASTAGES * ST_W_BYTES)
#define ST_SM_BAR (ST_SM_P + ST_PSTAGES * ST_P_BYTES)
#define ST_SM_TOTAL (ST_SM_BAR + 256)

__global__ void __launch_bounds__(ST_THREADS, 1)
    k_skin_tc(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_a,
              const __grid_constant__ CUtensorMap map_p, int B, int V, int Vp, int n_vt, int n_ch,
              float *__restrict__ verts) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + ST_SM_BAR;
  const uint32_t full_a = bar0 + 0, empty_a = bar0 + 16, full_w = bar0 + 32, empty_w = bar0 + 40;
  const uint32_t tmem_full = bar0 + 48, tmem_empty = bar0 + 64;
  const uint32_t full_p = bar0 + 128, empty_p = bar0 + 160;
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + ST_SM_BAR + 96);

  // (the shuffle tells the compiler the warp index is warp-uniform: role branches and the addresses
  // derived from it stay in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int total = n_vt * n_ch;
  const int t0 = (int)(((long long)blockIdx.x * total) / gridDim.x);
  const int t1 = (int)(((long long)(blockIdx.x + 1) * total) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < ST_ASTAGES; ++i) {
      mbar_init(full_a + 8 * i, 1);
      mbar_init(empty_a + 8 * i, 1);
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, 8);   // one arrival per epilogue warp
    }
    for (int i = 0; i < ST_PSTAGES; ++i) {
      mbar_init(full_p + 8 * i, 1);
      mbar_init(empty_p + 8 * i, 8);
    }
    mbar_init(full_w, 1);
    mbar_init(empty_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + ST_SM_BAR + 96), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      // tiles are ordered sample-chunk major: consecutive tiles of a CTA walk adjacent vertex
      // tiles of the same 16 samples (adjacent 512 B / 1536 B segments of the same v_posed /
      // verts rows).  The A16 chunk is loaded once per chunk, W16 tiles (L2-resident, 1.8 MB in
      // total) stream through a 2-deep ring.
      int cur_ch = -1, a_loads = 0, stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        int ch = t / n_vt, vt = t % n_vt;
        if (ch != cur_ch) {
          if (a_loads > 0) mbar_wait(empty_w, (a_loads - 1) & 1);
          mbar_expect_tx(full_w, ST_A_BYTES);
          tma_load_2d(sbase + ST_SM_A, &map_a, 0, ch * ST_N, full_w);
          ++a_loads;
          cur_ch = ch;
        }
        mbar_wait(empty_a + 8 * stage, phase ^ 1);
        mbar_expect_tx(full_a + 8 * stage, ST_W_BYTES);
        tma_load_2d(sbase + ST_SM_W + stage * ST_W_BYTES, &map_w, 0, vt * ST_VT, full_a + 8 * stage);
        if (++stage == ST_ASTAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // all 32 lanes run the loop, the elected lane issues (tc_ptx.cuh: elect_one)
    {
      constexpr uint32_t idesc = umma_idesc_f16(ST_VT, ST_N);
      const uint64_t desc_w0 = umma_desc_sw128(sbase + ST_SM_W), a_desc = umma_desc_sw128(sbase + ST_SM_A);
      int cur_ch = -1, a_loads = 0, stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = t0; t < t1; ++t) {
        int ch = t / n_vt;
        if (ch != cur_ch) {
          mbar_wait(full_w, a_loads & 1);
          ++a_loads;
          cur_ch = ch;
        }
        mbar_wait(tmem_empty + 8 * acc, acc_phase ^ 1);
        mbar_wait(full_a + 8 * stage, phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const uint64_t w_desc = umma_desc_add(desc_w0, stage * ST_W_BYTES);
        const bool last_of_ch = (t + 1 == t1) || ((t + 1) / n_vt != ch);
        if (elect_one()) {
          // five K = 16 steps, (W window, A window) as in the table at the top; a window is 32 B = 2 units
          tc_mma_f16(d_tmem, w_desc + 0, a_desc + 0, idesc, 0);
          tc_mma_f16(d_tmem, w_desc + 2, a_desc + 2, idesc, 1);
          tc_mma_f16(d_tmem, w_desc + 0, a_desc + 4, idesc, 1);
          tc_mma_f16(d_tmem, w_desc + 4, a_desc + 0, idesc, 1);
          tc_mma_f16(d_tmem, w_desc + 6, a_desc + 2, idesc, 1);
          tc_commit(empty_a + 8 * stage);
          tc_commit(tmem_full + 8 * acc);
          if (last_of_ch) tc_commit(empty_w);
        }
        __syncwarp();
        if (++stage == ST_ASTAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp == 10) {
    // =========================== v_posed tile producer ===========================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        int ch = t / n_vt, vt = t % n_vt;
        mbar_wait(empty_p + 8 * stage, phase ^ 1);
        mbar_expect_tx(full_p + 8 * stage, ST_P_BYTES);
        for (int cc = 0; cc < 3; ++cc)
          tma_load_2d(sbase + ST_SM_P + stage * ST_P_BYTES + cc * (ST_S * ST_VT * 4), &map_p, cc * Vp + vt * ST_VT,
                      ch * ST_S, full_p + 8 * stage);
        if (++stage == ST_PSTAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    // Two warps per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31); each takes
    // half of the tile's 16 samples, 4 samples at a time so the TMEM loads, the shared-memory
    // reads of v_posed and the stores of one group overlap instead of serialising.
    const int q = warp & 3;                       // vertices 32q .. 32q+31 of the tile
    const int half = (warp - 2) >> 2;             // sample groups {2*half, 2*half+1}
    int acc = 0, acc_phase = 0, pst = 0, pphase = 0;
    for (int t = t0; t < t1; ++t) {
      int ch = t / n_vt, vt = t % n_vt;
      int v0 = vt * ST_VT + 32 * q;               // first vertex of this warp
      int nflt = 3 * min(32, V - v0);             // floats of verts this warp owns per sample (<= 0: none)
      mbar_wait(full_p + 8 * pst, pphase);        // v_posed tile landed in shared memory
      float *ptile = reinterpret_cast<float *>(smem + ST_SM_P + pst * ST_P_BYTES) + 32 * q;
      mbar_wait(tmem_full + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16) + acc * 256;
#pragma unroll 1
      for (int g2 = 0; g2 < 2; ++g2) {
        const int sg = 2 * half + g2;
        uint32_t r[48];                           // T of 4 samples: 48 fp32 columns
        tc_ld_32x32(trow + sg * 48, r);
        tc_ld_32x16(trow + sg * 48 + 32, r + 32);
        float p[4][3];
#pragma unroll
        for (int si = 0; si < 4; ++si) {
          const float *pr = ptile + (sg * 4 + si) * ST_VT + lane;   // [xyz][sample][vertex]
          p[si][0] = pr[0];
          p[si][1] = pr[ST_S * ST_VT];
          p[si][2] = pr[2 * ST_S * ST_VT];
        }
        tc_wait_ld();
        if (g2 == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + 8 * acc);   // this warp has read its share of the accumulator
        }
        float o[4][3];
#pragma unroll
        for (int si = 0; si < 4; ++si) {
          const uint32_t *T = r + 12 * si;
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
            o[si][rr] = fmaf(__uint_as_float(T[4 * rr]), p[si][0],
                             fmaf(__uint_as_float(T[4 * rr + 1]), p[si][1],
                                  fmaf(__uint_as_float(T[4 * rr + 2]), p[si][2], __uint_as_float(T[4 * rr + 3]))));
        }
        // verts[b][v][xyz]: three strided scalar stores per sample.  The same LSU wavefronts as a
        // shared-memory transpose followed by coalesced stores, with a third of the instructions
        // (measured 134 -> 126 us); L2 merges the partial sectors.
#pragma unroll
        for (int si = 0; si < 4; ++si) {
          int b = ch * ST_S + sg * 4 + si;
          if (b >= B || 3 * lane >= nflt) continue;
          float *dst = verts + ((size_t)b * V + v0 + lane) * 3;
          __stcs(dst, o[si][0]);
          __stcs(dst + 1, o[si][1]);
          __stcs(dst + 2, o[si][2]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_p + 8 * pst);        // this warp is done with the v_posed tile
      if (++pst == ST_PSTAGES) {
        pst = 0;
        pphase ^= 1;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// W16[v][64]: the window layout at the top of this file; rows >= V are zero.
__global__ void k_build_w16(int V, int Vp, const float *__restrict__ W, __half *__restrict__ W16) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Vp * 64) return;
  int v = i / 64, k = i % 64;
  float out = 0.f;
  if (v < V && k < 56) {
    // joint and part (hi / lo) held by column k
    int j = k < 24 ? k : (k < 32 ? k - 8 : (k < 48 ? k - 32 : k - 32));
    bool lo = k >= 32;
    float w = W[(size_t)v * NJ + j];
    float hi = __half2float(__float2half_rn(w));
    out = lo ? (w - hi) : hi;
  }
  W16[i] = __float2half_rn(out);
}

int skin_tc_init(smplb_ctx *c) {
  c->skin_tc_ok = false;
  CUDA_TRY(cudaMalloc((void **)&c->d_W16, (size_t)c->Vp * 64 * sizeof(__half)));
  k_build_w16<<<cdiv(c->Vp * 64, 256), 256, 0, c->stream>>>(c->V, c->Vp, c->d_W, (__half *)c->d_W16);
  c->launches++;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaFuncSetAttribute(k_skin_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SM_TOTAL));
  TRY(tc_make_map(c->map_w, 0, c->d_W16, ST_KP, (uint64_t)c->Vp, ST_KP * 2, 64, ST_VT));
  if (!c->num_sms) CUDA_TRY(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device));
  c->skin_tc_ok = true;
  return 0;
}

// act: the same kernel on the active vertices (W16_act, v_posed_act) -> verts_act [B][n_act][3].
int launch_skin_fwd_tc(smplb_ctx *c, int B, const void *A16, const float *v_posed, float *verts, bool act) {
  RET_IF(!c->skin_tc_ok || (act && !c->compact_ok), SMPLB_ESTATE, "tcgen05 skinning path is not initialised");
  const int V = act ? c->n_act : c->V, Vp = act ? c->Vpa : c->Vp, pitch = act ? c->pitch_act : c->pitch;
  alignas(64) CUtensorMap map_a, map_p;
  TRY(tc_make_map(&map_a, 0, A16, ST_KP, (uint64_t)B * 12, ST_KP * 2, 64, ST_N));
  // v_posed [B][3 * Vp] fp32, box = 128 vertices x 16 samples of one coordinate plane, no swizzle
  TRY(tc_make_map(&map_p, 1, v_posed, (uint64_t)pitch, (uint64_t)B, (uint64_t)pitch * 4, ST_VT, ST_S, /*swizzle=*/0));
  int n_vt = Vp / ST_VT, n_ch = cdiv(B, ST_S);
  int total = n_vt * n_ch;
  int grid = total < c->num_sms ? total : c->num_sms;
  const CUtensorMap *mw = (const CUtensorMap *)(act ? c->map_w_act : c->map_w);
  LAUNCH(c, act ? "skin_fwd_tc_active" : "skin_fwd_tc", grid, ST_THREADS, ST_SM_TOTAL, k_skin_tc, *mw, map_a, map_p, B, V,
         Vp, n_vt, n_ch, verts);
  return 0;
}
