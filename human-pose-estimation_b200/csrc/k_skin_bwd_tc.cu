// Backward of linear blend skinning with BOTH of its contractions on the tensor cores (the dense backward of the
// mesh-loss step; TF autodiff through batch_smpl.py:139-149 in the reference).
//
//   g[s, v, :]   = d_verts[s, v, :] + sum_k joint_regressor[k, v] d_joints[s, k, :]
//   T[v, (s, e)] = sum_j W[v, j] A[s, j, e]                       (the forward's contraction, k_skin_tc.cu)
//   dp[s, v, :]  = T_R(s, v)^T g[s, v, :]                         (gradient w.r.t. v_posed)
//   dA[s, j, r, d] = sum_v W[v, j] g[s, v, r] [v_posed; 1][s, v, d]
//
// k_skin_bwd (FP32 CUDA cores) spends 216 + 288 FMAs per (vertex, sample) on the two sums over joints / vertices and
// one shared-memory load per 2-3 of them; here both run as tcgen05 MMAs around the same epilogue threads:
//
//   MMA 1  T (128 vertices x 96 = 8 samples x 12) = W16 tile . A16 chunk^T, the five split-precision K = 16 steps of
//          k_skin_tc.cu, two TMEM stages;
//   epilogue (thread = vertex = TMEM lane, two warps per lane quarter, 4 samples each): reads T, v_posed (TMA-staged
//          tile) and g, writes dp (as the bf16 hi | lo | hi operand row of the blend-transpose GEMM, or fp32), and forms
//          X[(s, r, d), v] = g_r [p; 1]_d as bf16 hi / lo straight into a K-major SWIZZLE_128B operand in shared memory
//          (bf16: the upstream gradient's magnitude is arbitrary, and fp16 would underflow; 16 significand bits);
//   MMA 2  dA^T ((s, r, d) x 32 joints, 24 used) += X . WT^T over the tile's 128 vertices: X_hi.W_hi + X_lo.W_hi +
//          X_hi.W_lo, 24 MMAs of 128 x 32 x 16, accumulated in TMEM over the vertex tiles of a work unit;
//   unit end: three epilogue warps read dA^T (lane = (s, r, d)) and write dA_part[split][s][j][4 r + d].
//
// A work unit = (chunk of 8 samples, one of VS vertex ranges); units go round-robin over the CTAs, the VS partials per
// sample are added in fixed order by the pose backward (deterministic, no float atomics).  Warps: 0 TMA producer of the
// MMA operands, 1 MMA issuer, 2-9 epilogue, 10 TMA producer of the v_posed tiles.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "smplb_internal.h"
#include "tc_ptx.cuh"

#define SBT_VT 128
#define SBT_S 8
#define SBT_N (12 * SBT_S)                     // 96
#define SBT_NJ 32                              // MMA 2's N: 24 joints padded
#define SBT_SPW 2                              // samples per epilogue warp and tile
#define SBT_EW (4 * SBT_S / SBT_SPW)           // epilogue warps: SBT_S / SBT_SPW per TMEM lane quarter
#define SBT_THREADS (32 * (3 + SBT_EW))
#define SBT_A_BYTES (SBT_N * 128)              // 12 KB: A16 rows of the chunk
#define SBT_W_BYTES (SBT_VT * 128)             // 16 KB: W16 tile
#define SBT_WT_PART (SBT_NJ * 128)             // 4 KB: 32 joints x 64 vertices of one (hi / lo, k-block)
#define SBT_WT_BYTES (4 * SBT_WT_PART)         // 16 KB per vertex tile
#define SBT_X_PART (128 * 128)                 // 16 KB: 128 rows (96 used) x 64 vertices of one (hi / lo, k-block)
#define SBT_X_BYTES (4 * SBT_X_PART)           // 64 KB
#define SBT_P_BYTES (3 * SBT_S * SBT_VT * 4)   // 12 KB: v_posed tile [xyz][8 samples][128 vertices]
#define SBT_WSTAGES 2
#define SBT_PSTAGES 3
#define SBT_SM_A 0
#define SBT_SM_W (SBT_SM_A + SBT_A_BYTES)
#define SBT_SM_WT (SBT_SM_W + SBT_WSTAGES * SBT_W_BYTES)
#define SBT_SM_X (SBT_SM_WT + SBT_WSTAGES * SBT_WT_BYTES)
#define SBT_SM_P (SBT_SM_X + SBT_X_BYTES)
#define SBT_SM_BAR (SBT_SM_P + SBT_PSTAGES * SBT_P_BYTES)
#define SBT_SM_TOTAL (SBT_SM_BAR + 256)
#define SBT_TCOL_D2 (2 * SBT_N)                // TMEM: T stage 0, T stage 1, dA^T (32 columns)

__global__ void __launch_bounds__(SBT_THREADS, 1)
    k_skin_bwd_tc(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_a,
                  const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_wt, int B, int V, int K,
                  int Vp, int n_vt, int n_ch, int VS, const float *__restrict__ d_verts, const float *__restrict__ d_joints,
                  const int *__restrict__ voff, const int *__restrict__ vk, const float *__restrict__ vval,
                  float *__restrict__ dp, __nv_bfloat16 *__restrict__ dp16, float *__restrict__ dA_part) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + SBT_SM_BAR;
  const uint32_t full_a = bar0 + 0, empty_a = bar0 + 8;            // A16 chunk
  const uint32_t full_w = bar0 + 16, empty_w = bar0 + 32;          // W16 ring (2)
  const uint32_t full_wt = bar0 + 48, empty_wt = bar0 + 64;        // WT ring (2)
  const uint32_t tmem_full = bar0 + 80, tmem_empty = bar0 + 96;    // T stages (2)
  const uint32_t x_full = bar0 + 112, x_empty = bar0 + 120;        // X operand (single buffer)
  const uint32_t d2_full = bar0 + 128, d2_empty = bar0 + 136;      // dA^T accumulator
  const uint32_t full_p = bar0 + 144, empty_p = bar0 + 176;        // v_posed ring (3)
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + SBT_SM_BAR + 224);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_units = n_ch * VS;
  // unit u: chunk u / VS, vertex tiles [vt0, vt1)
  auto unit_range = [&](int u, int &ch, int &sp, int &vt0, int &vt1) {
    ch = u / VS;
    sp = u % VS;
    vt0 = (int)(((long long)sp * n_vt) / VS);
    vt1 = (int)(((long long)(sp + 1) * n_vt) / VS);
  };

  if (threadIdx.x == 0) {
    mbar_init(full_a, 1);
    mbar_init(empty_a, 1);
    for (int i = 0; i < SBT_WSTAGES; ++i) {
      mbar_init(full_w + 8 * i, 1);
      mbar_init(empty_w + 8 * i, 1);
      mbar_init(full_wt + 8 * i, 1);
      mbar_init(empty_wt + 8 * i, 1);
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, SBT_EW);   // one arrival per epilogue warp
    }
    mbar_init(x_full, SBT_EW);
    mbar_init(x_empty, 1);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, SBT_EW);
    for (int i = 0; i < SBT_PSTAGES; ++i) {
      mbar_init(full_p + 8 * i, 1);
      mbar_init(empty_p + 8 * i, SBT_EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows 96..127 of the X operand are never written: zero them once (they only reach TMEM lanes nobody reads, but NaN
  // bit patterns in uninitialised shared memory must not enter the MMA)
  for (int i = threadIdx.x; i < SBT_X_BYTES / 16; i += SBT_THREADS)
    reinterpret_cast<uint4 *>(smem + SBT_SM_X)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SBT_SM_BAR + 224), "n"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =========================== TMA producer: A16 chunk per unit, W16 + WT tiles per vertex tile ===========================
    if (lane == 0) {
      int stage = 0, phase = 0, units = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++units) {
        int ch, sp, vt0, vt1;
        unit_range(u, ch, sp, vt0, vt1);
        if (units > 0) mbar_wait(empty_a, (units - 1) & 1);     // every MMA 1 of the previous unit has completed
        mbar_expect_tx(full_a, SBT_A_BYTES);
        tma_load_2d(sbase + SBT_SM_A, &map_a, 0, ch * SBT_N, full_a);
        for (int vt = vt0; vt < vt1; ++vt) {
          mbar_wait(empty_w + 8 * stage, phase ^ 1);
          mbar_expect_tx(full_w + 8 * stage, SBT_W_BYTES);
          tma_load_2d(sbase + SBT_SM_W + stage * SBT_W_BYTES, &map_w, 0, vt * SBT_VT, full_w + 8 * stage);
          mbar_wait(empty_wt + 8 * stage, phase ^ 1);
          mbar_expect_tx(full_wt + 8 * stage, SBT_WT_BYTES);
          for (int h = 0; h < 2; ++h)
            for (int kb = 0; kb < 2; ++kb)
              tma_load_2d(sbase + SBT_SM_WT + stage * SBT_WT_BYTES + (2 * h + kb) * SBT_WT_PART, &map_wt, vt * SBT_VT + 64 * kb,
                          32 * h, full_wt + 8 * stage);
          if (++stage == SBT_WSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (all lanes run the loop, the elected lane issues) ===========================
    constexpr uint32_t idesc1 = umma_idesc_f16(SBT_VT, SBT_N);
    constexpr uint32_t idesc2 = umma_idesc_f16(128, SBT_NJ) | (1u << 7) | (1u << 10);   // bf16 operands
    const uint64_t desc_a = umma_desc_sw128(sbase + SBT_SM_A), desc_w0 = umma_desc_sw128(sbase + SBT_SM_W);
    const uint64_t desc_x0 = umma_desc_sw128(sbase + SBT_SM_X), desc_wt0 = umma_desc_sw128(sbase + SBT_SM_WT);
    int stage1 = 0, phase1 = 0, acc = 0, acc_phase = 0;      // MMA 1 runs one tile ahead of MMA 2
    int stage2 = 0, phase2 = 0, xcount = 0, units = 0;
    auto mma1 = [&](bool last_of_unit) {
      mbar_wait(tmem_empty + 8 * acc, acc_phase ^ 1);
      mbar_wait(full_w + 8 * stage1, phase1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * SBT_N;
      const uint64_t w_desc = umma_desc_add(desc_w0, stage1 * SBT_W_BYTES);
      if (elect_one()) {
        // (W window, A window) pairs of the table in k_skin_tc.cu
        tc_mma_f16(d_tmem, w_desc + 0, desc_a + 0, idesc1, 0);
        tc_mma_f16(d_tmem, w_desc + 2, desc_a + 2, idesc1, 1);
        tc_mma_f16(d_tmem, w_desc + 0, desc_a + 4, idesc1, 1);
        tc_mma_f16(d_tmem, w_desc + 4, desc_a + 0, idesc1, 1);
        tc_mma_f16(d_tmem, w_desc + 6, desc_a + 2, idesc1, 1);
        tc_commit(empty_w + 8 * stage1);
        tc_commit(tmem_full + 8 * acc);
        if (last_of_unit) tc_commit(empty_a);
      }
      __syncwarp();
      if (++stage1 == SBT_WSTAGES) {
        stage1 = 0;
        phase1 ^= 1;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    };
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++units) {
      int ch, sp, vt0, vt1;
      unit_range(u, ch, sp, vt0, vt1);
      mbar_wait(full_a, units & 1);
      mbar_wait(d2_empty, (units & 1) ^ 1);     // the previous unit's dA^T has been read out
      tc_fence_after();
      mma1(vt0 + 1 == vt1);
      for (int vt = vt0; vt < vt1; ++vt) {
        if (vt + 1 < vt1) mma1(vt + 2 == vt1);
        // MMA 2 of tile vt: dA^T += X . WT^T
        mbar_wait(x_full, xcount & 1);
        mbar_wait(full_wt + 8 * stage2, phase2);
        tc_fence_after();
        const uint64_t wt_desc = umma_desc_add(desc_wt0, stage2 * SBT_WT_BYTES);
        if (elect_one()) {
          const uint32_t d2 = tmem_base + SBT_TCOL_D2;
          uint32_t accum = vt != vt0;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // parts: X [hi kb0, hi kb1, lo kb0, lo kb1] x 16 KB; WT [hi kb0, hi kb1, lo kb0, lo kb1] x 4 KB
              const uint64_t xh = desc_x0 + ((kb * SBT_X_PART) >> 4) + 2 * k, xl = xh + ((2 * SBT_X_PART) >> 4);
              const uint64_t wh = wt_desc + ((kb * SBT_WT_PART) >> 4) + 2 * k, wl = wh + ((2 * SBT_WT_PART) >> 4);
              tc_mma_f16(d2, xh, wh, idesc2, accum);
              tc_mma_f16(d2, xl, wh, idesc2, 1);
              tc_mma_f16(d2, xh, wl, idesc2, 1);
              accum = 1;
            }
          tc_commit(x_empty);
          tc_commit(empty_wt + 8 * stage2);
          if (vt + 1 == vt1) tc_commit(d2_full);
        }
        __syncwarp();
        ++xcount;
        if (++stage2 == SBT_WSTAGES) {
          stage2 = 0;
          phase2 ^= 1;
        }
      }
    }
  } else if (warp == 2 + SBT_EW) {
    // =========================== v_posed tile producer ===========================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        int ch, sp, vt0, vt1;
        unit_range(u, ch, sp, vt0, vt1);
        for (int vt = vt0; vt < vt1; ++vt) {
          mbar_wait(empty_p + 8 * stage, phase ^ 1);
          mbar_expect_tx(full_p + 8 * stage, SBT_P_BYTES);
          for (int cc = 0; cc < 3; ++cc)
            tma_load_2d(sbase + SBT_SM_P + stage * SBT_P_BYTES + cc * (SBT_S * SBT_VT * 4), &map_p, cc * Vp + vt * SBT_VT,
                        ch * SBT_S, full_p + 8 * stage);
          if (++stage == SBT_PSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    const int q = warp & 3;                       // vertices 32 q .. 32 q + 31 of the tile (TMEM lane quarter)
    const int half = (warp - 2) >> 2;             // samples SPW half .. SPW half + SPW - 1 of the chunk
    const int vl = 32 * q + lane;                 // vertex within the tile = column of the X operand
    // byte offset of column vl inside a 128-byte row of its k-block, before the swizzle XOR with the row
    const uint32_t x_part = sbase + SBT_SM_X + (vl >> 6) * SBT_X_PART;
    const uint32_t x_chunk = (uint32_t)((vl & 63) >> 3), x_in = (uint32_t)((vl & 7) * 2);
    int acc = 0, acc_phase = 0, pst = 0, pphase = 0, xcount = 0, units = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++units) {
      int ch, sp, vt0, vt1;
      unit_range(u, ch, sp, vt0, vt1);
      for (int vt = vt0; vt < vt1; ++vt) {
        const int v = vt * SBT_VT + vl;
        const bool v_ok = v < V;
        // the upstream gradient of this thread's (vertex, 4 samples) first: global loads that depend on nothing the
        // MMAs produce, issued before any wait (all twelve d_verts loads in flight together; the joint-regressor term
        // only exists for the few vertices with an entry, its row bounds are read once per tile)
        float g[SBT_SPW][3], p[SBT_SPW][3];
        int e0 = 0, e1 = 0;
        if (v_ok && d_joints) {
          e0 = voff[v];
          e1 = voff[v + 1];
        }
#pragma unroll
        for (int si = 0; si < SBT_SPW; ++si) {
          const int b = ch * SBT_S + SBT_SPW * half + si;
          g[si][0] = g[si][1] = g[si][2] = 0.f;
          if (v_ok && b < B && d_verts) {
            const float *dv = d_verts + ((size_t)b * V + v) * 3;
            g[si][0] = dv[0];
            g[si][1] = dv[1];
            g[si][2] = dv[2];
          }
        }
        for (int e = e0; e < e1; ++e) {
          const int k = vk[e];
          const float wv = vval[e];
#pragma unroll
          for (int si = 0; si < SBT_SPW; ++si) {
            const int b = ch * SBT_S + SBT_SPW * half + si;
            if (b < B) {
              const float *dj = d_joints + ((size_t)b * K + k) * 3;
              g[si][0] = fmaf(wv, dj[0], g[si][0]);
              g[si][1] = fmaf(wv, dj[1], g[si][1]);
              g[si][2] = fmaf(wv, dj[2], g[si][2]);
            }
          }
        }
        mbar_wait(tmem_full + 8 * acc, acc_phase);
        tc_fence_after();
        uint32_t r[12 * SBT_SPW];                 // T of this warp's samples
        const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16) + acc * SBT_N + half * (12 * SBT_SPW);
        if (SBT_SPW == 4) {
          tc_ld_32x32(trow, r);
          tc_ld_32x16(trow + 32, r + 32);
        } else {
          tc_ld_32x16(trow, r);
          tc_ld_32x8(trow + 16, r + 16);
        }
        mbar_wait(full_p + 8 * pst, pphase);
        const float *ptile = reinterpret_cast<const float *>(smem + SBT_SM_P + pst * SBT_P_BYTES) + vl;
#pragma unroll
        for (int si = 0; si < SBT_SPW; ++si) {
          const int sl = SBT_SPW * half + si, b = ch * SBT_S + sl;
          p[si][0] = p[si][1] = p[si][2] = 0.f;
          if (v_ok && b < B) {
            const float *pr = ptile + sl * SBT_VT;       // [xyz][sample][vertex]
            p[si][0] = pr[0];
            p[si][1] = pr[SBT_S * SBT_VT];
            p[si][2] = pr[2 * SBT_S * SBT_VT];
          }
        }
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(tmem_empty + 8 * acc);      // this warp has read its share of T
          mbar_arrive(empty_p + 8 * pst);         // ... and of the v_posed tile
        }
        // dp = T_R^T g
#pragma unroll
        for (int si = 0; si < SBT_SPW; ++si) {
          const int b = ch * SBT_S + SBT_SPW * half + si;
          if (!(v_ok && b < B)) continue;
          const uint32_t *T = r + 12 * si;
          float d[3];
#pragma unroll
          for (int c = 0; c < 3; ++c)
            d[c] = __uint_as_float(T[c]) * g[si][0] + __uint_as_float(T[4 + c]) * g[si][1] + __uint_as_float(T[8 + c]) * g[si][2];
          if (dp16) {
            // operand row of the tcgen05 blend-transpose GEMM: [hi | lo | hi], each 3 Vp wide (k_skin.cu)
            const size_t P3 = 3 * (size_t)Vp;
            __nv_bfloat16 *o = dp16 + (size_t)b * (3 * P3) + v;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const __nv_bfloat16 hi = __float2bfloat16_rn(d[c]);
              const __nv_bfloat16 lo = __float2bfloat16_rn(d[c] - __bfloat162float(hi));
              o[c * (size_t)Vp] = hi;
              o[P3 + c * (size_t)Vp] = lo;
              o[2 * P3 + c * (size_t)Vp] = hi;
            }
          } else {
            float *o = dp + (size_t)b * (3 * (size_t)Vp) + v;
            o[0] = d[0];
            o[Vp] = d[1];
            o[2 * (size_t)Vp] = d[2];
          }
        }
        // X[(s, r, d), v] = g_r [p; 1]_d -> bf16 hi / lo into the MMA 2 operand (the previous tile's MMA 2 must be done)
        mbar_wait(x_empty, (xcount & 1) ^ 1);
#pragma unroll
        for (int si = 0; si < SBT_SPW; ++si) {
          const float one = (v_ok && ch * SBT_S + SBT_SPW * half + si < B) ? 1.0f : 0.0f;
          const float ph[4] = {p[si][0], p[si][1], p[si][2], one};
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) {
              const float x = g[si][rr] * ph[dd];
              const __nv_bfloat16 hi = __float2bfloat16_rn(x);
              const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
              const uint32_t row = (uint32_t)((SBT_SPW * half + si) * 12 + 4 * rr + dd);
              const uint32_t addr = x_part + row * 128 + ((x_chunk ^ (row & 7)) << 4) + x_in;
              asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(__bfloat16_as_ushort(hi)) : "memory");
              asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr + 2 * SBT_X_PART), "h"(__bfloat16_as_ushort(lo)) : "memory");
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(x_full);
        ++xcount;
        if (++pst == SBT_PSTAGES) {
          pst = 0;
          pphase ^= 1;
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      // ---- unit end: dA^T (lane = (s, r, d), column = joint) -> dA_part[sp][b][j][4 r + d]
      mbar_wait(d2_full, units & 1);
      tc_fence_after();
      if (half == 0 && q < 3) {
        uint32_t dj[32];
        tc_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + SBT_TCOL_D2, dj);
        tc_wait_ld();
        const int m = 32 * q + lane, sl = m / 12, e = m % 12;
        const int b = ch * SBT_S + sl;
        if (b < B) {
          float *o = dA_part + (((size_t)sp * B + b) * NJ) * 12 + e;
#pragma unroll
          for (int j = 0; j < NJ; ++j) o[j * 12] = __uint_as_float(dj[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
  }
}

// WT[64][Vp] bf16: rows 0..23 = hi(W[v][j]), rows 32..55 = lo, the others zero (MMA 2's N is 32)
__global__ void k_build_wt16(int V, int Vp, const float *__restrict__ W, __nv_bfloat16 *__restrict__ WT) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * Vp) return;
  int row = i / Vp, v = i % Vp;
  int j = row & 31;
  float out = 0.f;
  if (v < V && j < NJ) {
    float w = W[(size_t)v * NJ + j];
    float hi = __bfloat162float(__float2bfloat16_rn(w));
    out = row >= 32 ? (w - hi) : hi;
  }
  WT[i] = __float2bfloat16_rn(out);
}

int skin_bwd_tc_init(smplb_ctx *c) {
  c->skin_bwd_tc_ok = false;
  if (!c->skin_tc_ok) return 0;
  CUDA_TRY(cudaMalloc((void **)&c->d_WT16, (size_t)64 * c->Vp * sizeof(__nv_bfloat16)));
  k_build_wt16<<<cdiv(64 * c->Vp, 256), 256, 0, c->stream>>>(c->V, c->Vp, c->d_W, (__nv_bfloat16 *)c->d_WT16);
  c->launches++;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaFuncSetAttribute(k_skin_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SBT_SM_TOTAL));
  // (bf16 and fp16 are both 2-byte elements: the tensor map only moves bytes)
  TRY(tc_make_map(c->map_wt, 0, c->d_WT16, (uint64_t)c->Vp, 64, (uint64_t)c->Vp * 2, 64, SBT_NJ));
  c->skin_bwd_tc_ok = true;
  return 0;
}

// dense walk only (every vertex, upstream d_verts and / or d_joints); *n_parts = partial sums written per sample
int launch_skin_bwd_tc(smplb_ctx *c, int B, const void *A16, const float *v_posed, const float *d_verts, const float *d_joints,
                       float *dp, void *dp16, float *dA_part, int *n_parts) {
  RET_IF(!c->skin_bwd_tc_ok, SMPLB_ESTATE, "tcgen05 skinning backward is not initialised");
  alignas(64) CUtensorMap map_a, map_p;
  TRY(tc_make_map(&map_a, 0, A16, 64, (uint64_t)B * 12, 128, 64, SBT_N));
  TRY(tc_make_map(&map_p, 1, v_posed, (uint64_t)c->pitch, (uint64_t)B, (uint64_t)c->pitch * 4, SBT_VT, SBT_S, /*swizzle=*/0));
  const int n_vt = c->Vp / SBT_VT, n_ch = cdiv(B, SBT_S);
  // Vertex ranges per chunk (= partial sums per sample, at most skin_bwd_splits(B)): units go round-robin over the SMs,
  // so pick the split whose last round is fullest (B = 1024: 8 splits = 1024 units = 6.9 rounds; 5 splits would be
  // 4.3 rounds, the SMs with a fifth unit finishing 25 % after the others), with at least 3 vertex tiles per unit.
  int VS = 1;
  double best_eff = 0.0;
  for (int vs = 1; vs <= skin_bwd_splits(B) && vs * 3 <= n_vt; ++vs) {
    const int units = n_ch * vs;
    const double eff = (double)units / ((double)cdiv(units, c->num_sms) * c->num_sms);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      VS = vs;
    }
  }
  const int n_units = n_ch * VS;
  const int grid = n_units < c->num_sms ? n_units : c->num_sms;
  LAUNCH(c, "skin_bwd_tc", grid, SBT_THREADS, SBT_SM_TOTAL, k_skin_bwd_tc, *(const CUtensorMap *)c->map_w, map_a, map_p,
         *(const CUtensorMap *)c->map_wt, B, c->V, c->K, c->Vp, n_vt, n_ch, VS, d_verts, d_joints, c->d_vcsr_off, c->d_vcsr_k,
         c->d_vcsr_val, dp, (__nv_bfloat16 *)dp16, dA_part);
  *n_parts = VS;
  return 0;
}
