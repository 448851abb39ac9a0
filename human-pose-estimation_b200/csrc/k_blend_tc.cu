// Blend-shape contraction on the 5th-generation tensor cores (tcgen05 + TMEM, TMA-fed).
//
//   v_posed[b, n] = 2^-s * sum_k x16[b,k] * Dt16[n,k]          (batch_smpl.py:110-112, :126-132)
//
// Operands are fp16 (same 11-bit significand as tf32, at twice the rate) with fp32
// accumulation in TMEM.  The K axis (256 = 4 swizzle atoms of 64) carries
//   k   0..206  pose_feature            x posedirs * 2^s
//   k 207..216  beta_hi                 x shapedirs_hi * 2^s     } split-precision terms: the
//   k 217..226  beta_hi                 x shapedirs_lo * 2^s     } shape blend and the template
//   k 227..236  beta_lo                 x shapedirs_hi * 2^s     } are ~1 m and must keep fp32
//   k 237..239  1, 1, 1                 x v_template hi/mid/lo   } accuracy
//   k 240..255  0
// 2^s is a power of two chosen at create time so the constants sit in fp16's normal range;
// the epilogue multiplies by 2^-s (exact).
//
// Kernel shape: persistent, one CTA per SM, warp-specialised:
//   warp 0   TMA producer (x16 tile per sample block, Dt16 tiles through a 2-deep ring)
//   warp 1   tcgen05.mma issuer (one lane), TMEM allocator
//   warps 2-9  epilogue: tcgen05.ld -> scale -> swizzled smem staging -> TMA store
// Tile 256 (samples, two MMA sub-blocks) x 128 (vertex coordinates); two TMEM accumulator
// stages so the MMAs of tile i+1 overlap the epilogue of tile i.  The kernel is bound by its
// fp32 output stream and the L2 operand traffic that feeds it.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

#include "smplb_internal.h"

#define TC_BM 128            // MMA M (one sub-block of samples)
#define TC_MSUB 1            // sub-blocks per output tile: a CTA tile is 256 samples x 128 coordinates
#define TC_BN 128
#define TC_KP 256            // padded K
#define TC_KB 64             // K elements per 128-byte swizzle atom
#define TC_NKB (TC_KP / TC_KB)
#define TC_ASTAGES 4         // ring stages, one Dt16 k-block (16 KB) each
#define TC_THREADS 320

#define TILE_KB_BYTES (128 * 128)                       // one k-block of an operand tile: 128 rows x 128 B
#define SM_X_OFF 0                                      // resident x16 tiles: TC_MSUB x 4 k-blocks = 128 KB
#define SM_A_OFF (TC_MSUB * TC_NKB * TILE_KB_BYTES)     // Dt16 ring: 4 x 16 KB
#define SM_C_OFF (SM_A_OFF + TC_ASTAGES * TILE_KB_BYTES)  // epilogue staging: 8 warps x 4 KB
#define SM_BAR_OFF (SM_C_OFF + 8 * 4096)
#define SM_TOTAL (SM_BAR_OFF + 128)

#include "tc_ptx.cuh"

#define TC_IDESC umma_idesc_f16(TC_BM, TC_BN)

// Both streaming kernels of the forward turned out to be bound by the L2 <-> SM fabric
// (~8.3 TB/s of combined operand reads and output writes), not by HBM: with a 128-sample tile
// every 64 KB of output costs a 64 KB Dt16 operand tile from L2.  A CTA therefore owns two
// 128-sample sub-blocks (two resident x16 tiles, two accumulators) per Dt16 tile, halving
// the operand traffic per output byte.
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_blend_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_d,
               const __grid_constant__ CUtensorMap map_c, int n_mblk, int n_nblk, float inv_scale) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + SM_BAR_OFF;
  // barrier slots (8 bytes each)
  const uint32_t full_a = bar0 + 0, empty_a = bar0 + 32, full_x = bar0 + 64, empty_x = bar0 + 72;
  const uint32_t tmem_full = bar0 + 80, tmem_empty = bar0 + 96;
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + SM_BAR_OFF + 120);

  // (the shuffle tells the compiler the warp index is warp-uniform: role branches and the addresses
  // derived from it stay in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int total = n_mblk * n_nblk;          // n_mblk counts 256-sample super blocks
  const int t0 = (int)(((long long)blockIdx.x * total) / gridDim.x);
  const int t1 = (int)(((long long)(blockIdx.x + 1) * total) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_ASTAGES; ++i) {
      mbar_init(full_a + 8 * i, 1);
      mbar_init(empty_a + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, 8);   // one arrival per epilogue warp
    }
    mbar_init(full_x, 1);
    mbar_init(empty_x, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SM_BAR_OFF + 120), "n"(512)
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
      // tiles are ordered sample-block major: consecutive tiles of a CTA write adjacent 512 B
      // column ranges of the same 256 v_posed rows.  The two x16 tiles are loaded once per
      // sample block; Dt16 k-blocks (10.6 MB in total, L2-resident) stream through the ring.
      int cur_m = -1, x_loads = 0, stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        int m = t / n_nblk, n = t % n_nblk;
        if (m != cur_m) {
          if (x_loads > 0) mbar_wait(empty_x, (x_loads - 1) & 1);   // MMAs of the previous sample block retired
          mbar_expect_tx(full_x, TC_MSUB * TC_NKB * TILE_KB_BYTES);
          for (int u = 0; u < TC_MSUB; ++u)
            for (int kb = 0; kb < TC_NKB; ++kb)
              tma_load_2d(sbase + SM_X_OFF + (u * TC_NKB + kb) * TILE_KB_BYTES, &map_x, kb * TC_KB,
                          (m * TC_MSUB + u) * TC_BM, full_x);
          ++x_loads;
          cur_m = m;
        }
        for (int kb = 0; kb < TC_NKB; ++kb) {
          mbar_wait(empty_a + 8 * stage, phase ^ 1);
          mbar_expect_tx(full_a + 8 * stage, TILE_KB_BYTES);
          tma_load_2d(sbase + SM_A_OFF + stage * TILE_KB_BYTES, &map_d, kb * TC_KB, n * TC_BN, full_a + 8 * stage);
          if (++stage == TC_ASTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // all 32 lanes run the loop, the elected lane issues (tc_ptx.cuh: elect_one)
    {
      const uint64_t desc_x0 = umma_desc_sw128(sbase + SM_X_OFF), desc_a0 = umma_desc_sw128(sbase + SM_A_OFF);
      int cur_m = -1, x_loads = 0, stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = t0; t < t1; ++t) {
        int m = t / n_nblk;
        if (m != cur_m) {
          mbar_wait(full_x, x_loads & 1);
          ++x_loads;
          cur_m = m;
        }
        mbar_wait(tmem_empty + 8 * acc, acc_phase ^ 1);   // epilogue drained this accumulator pair
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < TC_NKB; ++kb) {
          mbar_wait(full_a + 8 * stage, phase);           // Dt16 k-block landed
          tc_fence_after();
          const uint64_t b_desc = umma_desc_add(desc_a0, stage * TILE_KB_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int u = 0; u < TC_MSUB; ++u) {
              // MMA "A" (M = samples) is a resident x16 tile, "B" (N = coordinates) the ring stage
              const uint64_t a_desc = umma_desc_add(desc_x0, (u * TC_NKB + kb) * TILE_KB_BYTES);
              const uint32_t d_tmem = tmem_base + (acc * TC_MSUB + u) * TC_BN;
#pragma unroll
              for (int k = 0; k < TC_KB / 16; ++k)   // advance 16 fp16 = 32 B inside the 128 B swizzle atom
                tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, TC_IDESC, (kb | k) != 0);
            }
            tc_commit(empty_a + 8 * stage);    // ring stage reusable once these MMAs retire
          }
          __syncwarp();
          if (++stage == TC_ASTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        const bool last_of_m = (t + 1 == t1) || ((t + 1) / n_nblk != m);
        if (elect_one()) {
          tc_commit(tmem_full + 8 * acc);      // accumulators ready for the epilogue
          if (last_of_m) tc_commit(empty_x);
        }
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    // Warp (id % 4) owns TMEM lanes 32*(id%4).. (the only lanes it may access).  With one
    // sub-block the two warps of a lane quarter split the tile's four 32-column chunks; with two
    // sub-blocks each takes one sub-block.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int u = TC_MSUB == 2 ? half : 0;
    const int chunk0 = TC_MSUB == 2 ? 0 : 2 * half;
    constexpr int NPAIR = TC_MSUB == 2 ? 2 : 1;     // pairs of 32-column chunks this warp drains per tile
    const uint32_t stage_base = sbase + SM_C_OFF + (warp - 2) * 4096;
    int acc = 0, acc_phase = 0;
    for (int t = t0; t < t1; ++t) {
      int m = t / n_nblk, n = t % n_nblk;
      mbar_wait(tmem_full + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(32 * q) << 16) + (acc * TC_MSUB + u) * TC_BN + chunk0 * 32;
#pragma unroll 1
      for (int pair = 0; pair < NPAIR; ++pair) {
        uint32_t r0[32], r1[32];
        tc_ld_32x32(tbase + pair * 64, r0);
        tc_ld_32x32(tbase + pair * 64 + 32, r1);
        tc_wait_ld();
        if (pair == NPAIR - 1) {
          // all of this warp's share of the accumulator is in registers: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + 8 * acc);
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const uint32_t *r = cc == 0 ? r0 : r1;
          if (lane == 0) tma_wait_read<0>();          // the staging buffer was read by this warp's previous store
          __syncwarp();
          uint32_t row_addr = stage_base + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v;
            v.x = __uint_as_float(r[4 * j + 0]) * inv_scale;
            v.y = __uint_as_float(r[4 * j + 1]) * inv_scale;
            v.z = __uint_as_float(r[4 * j + 2]) * inv_scale;
            v.w = __uint_as_float(r[4 * j + 3]) * inv_scale;
            uint32_t addr = row_addr + ((j ^ (lane & 7)) << 4);   // SWIZZLE_128B: 16B chunk ^= row % 8
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_c, stage_base, n * TC_BN + (chunk0 + pair * 2 + cc) * 32, (m * TC_MSUB + u) * TC_BM + 32 * q);
            tma_commit();
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) tma_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------- operand construction
__global__ void k_absmax(size_t n, const float *__restrict__ x, unsigned int *__restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order as uints
}

// Dt16[n][k] (row pitch 256 halves) from Dext fp32 [KX][pitch] (same planar column order n,
// zero in the padding columns); see the K map at the top.
__global__ void k_build_dt16(int pitch, int NB, float scale, const float *__restrict__ Dext,
                             __half *__restrict__ Dt) {
  int n = blockIdx.x * 4 + (threadIdx.x >> 6);
  int k0 = (threadIdx.x & 63) * 4;
  if (n >= pitch) return;
  for (int k = k0; k < k0 + 4; ++k) {
    float v = 0.f;
    {
      if (k < NPF) {
        v = Dext[(size_t)k * pitch + n] * scale;
      } else if (k < NPF + 3 * 10) {
        int slot = (k - NPF) / 10, bi = (k - NPF) % 10;
        if (bi < NB) {
          float s = Dext[(size_t)(NPF + bi) * pitch + n] * scale;
          float hi = __half2float(__float2half_rn(s));
          v = (slot == 1) ? (s - hi) : hi;    // slots: hi, lo, hi
        }
      } else if (k < NPF + 33) {
        float t = Dext[(size_t)(NPF + NB) * pitch + n] * scale;
        float hi = __half2float(__float2half_rn(t));
        float mid = __half2float(__float2half_rn(t - hi));
        int w = k - (NPF + 30);
        v = w == 0 ? hi : (w == 1 ? mid : (t - hi - mid));
      }
    }
    Dt[(size_t)n * TC_KP + k] = __float2half_rn(v);
  }
}

// ------------------------------------------------------------------------------ host side
// Called from smplb_create: builds Dt16 and picks the power-of-two scale.
int blend_tc_init(smplb_ctx *c) {
  c->tc_ok = false;
  if (c->NB > 10) return 0;   // the K map above reserves 10 slots per shape term
  unsigned int *d_max = nullptr;
  CUDA_TRY(cudaMalloc((void **)&d_max, 4));
  CUDA_TRY(cudaMemsetAsync(d_max, 0, 4, c->stream));
  size_t n = (size_t)(NPF + c->NB + 1) * c->pitch;
  k_absmax<<<296, 256, 0, c->stream>>>(n, c->d_Dext, d_max);
  unsigned int bits = 0;
  CUDA_TRY(cudaMemcpyAsync(&bits, d_max, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(d_max);
  float mx;
  memcpy(&mx, &bits, 4);
  int s = 0;
  if (mx > 0.f && isfinite(mx)) s = (int)floorf(log2f(16384.0f / mx));
  s = s < -8 ? -8 : (s > 24 ? 24 : s);
  c->tc_scale = ldexpf(1.0f, s);
  c->tc_inv_scale = ldexpf(1.0f, -s);
  CUDA_TRY(cudaMalloc((void **)&c->d_Dt16, (size_t)c->pitch * TC_KP * sizeof(__half)));
  k_build_dt16<<<cdiv(c->pitch, 4), 256, 0, c->stream>>>(c->pitch, c->NB, c->tc_scale, c->d_Dext,
                                                         (__half *)c->d_Dt16);
  c->launches += 2;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaFuncSetAttribute(k_blend_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  c->num_sms = sms;
  TRY(tc_make_map(c->map_d, 0, c->d_Dt16, TC_KP, (uint64_t)c->pitch,
                  TC_KP * 2, TC_KB, TC_BN));
  c->tc_ok = true;
  return 0;
}

// act: the same GEMM restricted to the active vertices (Dt16_act rows) -> v_posed_act [B][3*Vpa].
int launch_blend_fwd_tc(smplb_ctx *c, int B, const void *x16, float *v_posed, bool act) {
  RET_IF(!c->tc_ok || (act && !c->compact_ok), SMPLB_ESTATE, "tcgen05 blend path is not initialised");
  int pitch = act ? c->pitch_act : c->pitch;
  alignas(64) CUtensorMap map_x, map_c;
  TRY(tc_make_map(&map_x, 0, x16, TC_KP, (uint64_t)B, TC_KP * 2, TC_KB, TC_BM));
  TRY(tc_make_map(&map_c, 1, v_posed, (uint64_t)pitch, (uint64_t)B,
                  (uint64_t)pitch * 4, 32, 32));
  int n_mblk = cdiv(B, TC_BM * TC_MSUB), n_nblk = pitch / TC_BN;
  int total = n_mblk * n_nblk;
  int grid = total < c->num_sms ? total : c->num_sms;
  const CUtensorMap *md = (const CUtensorMap *)(act ? c->map_d_act : c->map_d);
  LAUNCH(c, act ? "blend_fwd_tc_active" : "blend_fwd_tc", grid, TC_THREADS, SM_TOTAL, k_blend_tc, map_x, *md, map_c, n_mblk,
         n_nblk, c->tc_inv_scale);
  return 0;
}

// Rows of a [rows][row_halves] fp16 matrix gathered by planar active index:
// out[(cc * Vpa + a)][:] = in[(cc * Vp + act[a])][:] for cc < ncoord, zero for a >= n_act.
__global__ void k_gather_rows16(int n_act, int Vpa, int Vp, int ncoord, int row_halves, const int *__restrict__ act,
                                const __half *__restrict__ in, __half *__restrict__ out) {
  int row = blockIdx.x;
  int cc = row / Vpa, a = row % Vpa;
  if (cc >= ncoord) return;
  for (int k = threadIdx.x; k < row_halves; k += blockDim.x)
    out[(size_t)row * row_halves + k] =
        a < n_act ? in[((size_t)cc * Vp + act[a]) * row_halves + k] : __float2half_rn(0.f);
}

int compact_tc_init(smplb_ctx *c) {
  c->compact_ok = false;
  if (!c->tc_ok || !c->skin_tc_ok || c->n_act >= c->V || c->n_act == 0) return 0;
  CUDA_TRY(cudaMalloc((void **)&c->d_Dt16_act, (size_t)c->pitch_act * TC_KP * sizeof(__half)));
  CUDA_TRY(cudaMalloc((void **)&c->d_W16_act, (size_t)c->Vpa * 64 * sizeof(__half)));
  k_gather_rows16<<<c->pitch_act, 128, 0, c->stream>>>(c->n_act, c->Vpa, c->Vp, 3, TC_KP, c->d_act_idx,
                                                      (const __half *)c->d_Dt16, (__half *)c->d_Dt16_act);
  k_gather_rows16<<<c->Vpa, 64, 0, c->stream>>>(c->n_act, c->Vpa, c->Vp, 1, 64, c->d_act_idx,
                                                (const __half *)c->d_W16, (__half *)c->d_W16_act);
  c->launches += 2;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  TRY(tc_make_map(c->map_d_act, 0, c->d_Dt16_act, TC_KP,
                  (uint64_t)c->pitch_act, TC_KP * 2, TC_KB, TC_BN));
  TRY(tc_make_map(c->map_w_act, 0, c->d_W16_act, 64, (uint64_t)c->Vpa,
                  64 * 2, 64, 128));
  c->compact_ok = true;
  return 0;
}
