// Fused blend shapes + linear blend skinning on CTA pairs (tcgen05 cta_group::2): the kernel of
// k_body_tc.cu with every MMA spanning the two SMs of a pair (batch_smpl.py:110-112, :126-132,
// :139-149).
//
// k_body_tc is bound by the bytes an SM exchanges with L2, and 40 % of what it reads are the "B"
// operands of its two contractions: the x16 tile (pose features / betas of NS samples) and the A16
// rows (the 3x4 joint transforms of the same samples).  Two CTAs that work on the same samples but
// on different vertex tiles need the same B operands.  A cta_group::2 MMA is built for exactly
// that: M = 256 = the pair's two vertex tiles (each CTA supplies its own 128 rows of "A" -- its
// Dt16 k-block or W16 tile -- and receives its own 128 accumulator lanes), N = NS or 12 * ST as
// before, and each CTA supplies only HALF of the N rows of "B" from its shared memory.  Per
// super-tile a CTA therefore loads 24 KB instead of 48 KB of x16 (per sample block) and 72 KB
// instead of 144 KB of A16.
//
// Roles per CTA as in k_body_tc: warp 0 blend-operand producer, warp 2 skinning-operand producer
// (both CTAs load their own halves; every complete_tx goes to the LEADER's full barrier, which
// expects the pair's bytes), warps 1 / 3 MMA issuers (leader only; tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to the barrier at the same offset in both CTAs),
// warps 4-11 epilogue on the CTA's own tensor memory (their "accumulator read" arrivals go to the
// leader's barriers, through the cluster's shared-memory window from the peer).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>

#include "smplb_internal.h"
#include "tc_ptx.cuh"
#include "body_common.cuh"

// -DFB_TIMING: per-role wait / busy cycle counts of two pairs (tools/fused_timing.py)
#ifdef FB_TIMING
#define TCLK() clock64()
#define TADD(acc, t) acc += clock64() - (t)
#else
#define TCLK() 0ll
#define TADD(acc, t)
#endif

template <int NS_, int ST_, int DSTAGES_, int ASTAGES_, int PRE_>
struct PairCfg {
  static constexpr int NS = NS_, ST = ST_, TBUF = 2, DSTAGES = DSTAGES_, ASTAGES = ASTAGES_, PRE = PRE_, EW = 2;
  static constexpr int TN = 12 * ST;                 // skinning MMA N
  static constexpr int NT = NS / ST;                 // skinning tiles per super-tile
  static constexpr int X_KB_HALF = (NS / 2) * 128;   // this CTA's rows of one k-block of the x16 tile
  static constexpr int X_HALF = 4 * X_KB_HALF;
  static constexpr int A_HALF = (TN / 2) * 128;      // this CTA's A16 rows of one skinning tile
  static constexpr int SM_X = 0;
  static constexpr int SM_D = SM_X + X_HALF;
  static constexpr int SM_W = SM_D + DSTAGES * FB_D_BYTES;
  static constexpr int SM_A = SM_W + 2 * FB_W_BYTES;
  static constexpr int SM_BAR = SM_A + ASTAGES * A_HALF;
  static constexpr int SM_TOTAL = SM_BAR + 512;
  static constexpr int TCOL = 3 * NS;                // first TMEM column of the T stages
  static constexpr int HS = ST / EW;                 // samples per epilogue warp and tile
  static constexpr int THREADS = 32 * (4 + 4 * EW);
  static_assert(NT % TBUF == 0 && PRE < NT, "tile counts");
  static_assert(HS == 4, "the epilogue is written for 4 samples per warp");
  static_assert(NS % 16 == 0 && TN % 16 == 0 && NS % ST == 0, "cta_group::2 MMAs take N in steps of 16");
  static_assert(3 * NS + TBUF * TN <= 512, "TMEM budget");
  static_assert(X_KB_HALF % 1024 == 0 && A_HALF % 1024 == 0, "swizzle atoms need 1024 B alignment");
  static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
  static_assert(DSTAGES <= 8 && ASTAGES <= 4, "barrier slots");
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
    k_body_pair(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_d,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_a, int B, int V, int Vp,
                int n_vp, int n_m, float inv_scale, float *__restrict__ verts) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + C::SM_BAR;
  const uint32_t full_x = bar0 + 0, empty_x = bar0 + 8, p_full = bar0 + 16, p_empty = bar0 + 24;
  const uint32_t full_d = bar0 + 32, empty_d = bar0 + 96;     // C::DSTAGES (<= 8) each
  const uint32_t full_w = bar0 + 160, empty_w = bar0 + 176;   // 2 each
  const uint32_t full_a = bar0 + 192, empty_a = bar0 + 224;   // C::ASTAGES (<= 4) each
  const uint32_t t_full = bar0 + 256, t_empty = bar0 + 272;   // C::TBUF (2) each
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + C::SM_BAR + 288);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;
  // Super-tile t of this pair -> sample block t / n_vp, vertex-tile pair t % n_vp; this CTA's
  // vertex tile is 2 * (t % n_vp) + crank.
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int total = n_vp * n_m;
  const int t0 = (int)(((long long)pair * total) / n_pairs);
  const int t1 = (int)(((long long)(pair + 1) * total) / n_pairs);

  if (threadIdx.x == 0) {
    // (the full_*, p_empty and t_empty barriers are only used in the leader)
    mbar_init(full_x, 1);
    mbar_init(empty_x, 1);
    mbar_init(p_full, 1);
    mbar_init(p_empty, 8 * C::EW);          // one arrival per epilogue warp of the pair
    for (int i = 0; i < C::DSTAGES; ++i) {
      mbar_init(full_d + 8 * i, 1);
      mbar_init(empty_d + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(full_w + 8 * i, 1);
      mbar_init(empty_w + 8 * i, 1);
    }
    for (int i = 0; i < C::ASTAGES; ++i) {
      mbar_init(full_a + 8 * i, 1);
      mbar_init(empty_a + 8 * i, 1);
    }
    for (int i = 0; i < C::TBUF; ++i) {
      mbar_init(t_full + 8 * i, 1);
      mbar_init(t_empty + 8 * i, 8 * C::EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // both CTAs of the pair allocate (same warp index, same destination offset)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C::SM_BAR + 288), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // the peer's barriers and tensor memory exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =========================== blend operand producer (both CTAs) ===========================
    if (lane == 0) {
      const uint32_t l_full_x = cluster_map_shared(full_x, 0), l_full_d = cluster_map_shared(full_d, 0);
      int cur_m = -1, x_loads = 0, stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int m = t / n_vp, vt = 2 * (t % n_vp) + crank;
        if (m != cur_m) {
          if (x_loads > 0) mbar_wait(empty_x, (x_loads - 1) & 1);
          if (leader) mbar_expect_tx(full_x, 2 * C::X_HALF);
          for (int kb = 0; kb < 4; ++kb)
            tma_load_2d_pair(sbase + C::SM_X + kb * C::X_KB_HALF, &map_x, kb * 64, m * C::NS + crank * (C::NS / 2), l_full_x);
          ++x_loads;
          cur_m = m;
        }
        for (int cc = 0; cc < 3; ++cc)
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(empty_d + 8 * stage, phase ^ 1);
            if (leader) mbar_expect_tx(full_d + 8 * stage, 2 * FB_D_BYTES);
            tma_load_2d_pair(sbase + C::SM_D + stage * FB_D_BYTES, &map_d, kb * 64, cc * Vp + vt * FB_VT, l_full_d + 8 * stage);
            if (++stage == C::DSTAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
      }
      // the leader's last "stage free" arrivals must have landed before this CTA may exit
      if (x_loads > 0) mbar_wait(empty_x, (x_loads - 1) & 1);
      for (int i = 0; i < C::DSTAGES; ++i) {
        mbar_wait(empty_d + 8 * stage, phase ^ 1);
        if (++stage == C::DSTAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // =========================== skinning operand producer (both CTAs) ===========================
    if (lane == 0) {
      const uint32_t l_full_w = cluster_map_shared(full_w, 0), l_full_a = cluster_map_shared(full_a, 0);
      int wbuf = 0, wphase = 0, stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int m = t / n_vp, vt = 2 * (t % n_vp) + crank;
        mbar_wait(empty_w + 8 * wbuf, wphase ^ 1);
        if (leader) mbar_expect_tx(full_w + 8 * wbuf, 2 * FB_W_BYTES);
        tma_load_2d_pair(sbase + C::SM_W + wbuf * FB_W_BYTES, &map_w, 0, vt * FB_VT, l_full_w + 8 * wbuf);
        if (++wbuf == 2) {
          wbuf = 0;
          wphase ^= 1;
        }
        for (int st = 0; st < C::NT; ++st) {
          mbar_wait(empty_a + 8 * stage, phase ^ 1);
          if (leader) mbar_expect_tx(full_a + 8 * stage, 2 * C::A_HALF);
          tma_load_2d_pair(sbase + C::SM_A + stage * C::A_HALF, &map_a, 0, (m * C::NS + st * C::ST) * 12 + crank * (C::TN / 2),
                           l_full_a + 8 * stage);
          if (++stage == C::ASTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      for (int i = 0; i < 2; ++i) {
        mbar_wait(empty_w + 8 * wbuf, wphase ^ 1);
        if (++wbuf == 2) {
          wbuf = 0;
          wphase ^= 1;
        }
      }
      for (int i = 0; i < C::ASTAGES; ++i) {
        mbar_wait(empty_a + 8 * stage, phase ^ 1);
        if (++stage == C::ASTAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // =========================== blend MMA issuer (leader) ===========================
    if (leader) {
      constexpr uint32_t idesc_p = umma_idesc_f16(2 * FB_VT, C::NS);
      const uint64_t desc_d0 = umma_desc_sw128(sbase + C::SM_D), desc_x0 = umma_desc_sw128(sbase + C::SM_X);
      int cur_m = -1, x_loads = 0, dstage = 0, dphase = 0, n_tiles = 0;
      [[maybe_unused]] long long w_pe = 0, w_fd = 0, w_tot = TCLK(), tq;
      for (int t = t0; t < t1; ++t, ++n_tiles) {
        const int m = t / n_vp;
        if (m != cur_m) {
          mbar_wait(full_x, x_loads & 1);
          ++x_loads;
          cur_m = m;
        }
        tq = TCLK();
        mbar_wait(p_empty, (n_tiles & 1) ^ 1);       // both epilogues have read the previous P
        TADD(w_pe, tq);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 3; ++cc) {
          const uint32_t d_tmem = tmem_base + cc * C::NS;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            tq = TCLK();
            mbar_wait(full_d + 8 * dstage, dphase);
            TADD(w_fd, tq);
            tc_fence_after();
            const uint64_t a_desc = umma_desc_add(desc_d0, dstage * FB_D_BYTES);
            const uint64_t b_desc = umma_desc_add(desc_x0, kb * C::X_KB_HALF);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kb == 3 && k == 3) continue;      // K = 240: the last 16 columns are zero padding
                tc_mma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_p, (kb | k) != 0);
              }
              tc_commit_pair(empty_d + 8 * dstage);
            }
            __syncwarp();
            if (++dstage == C::DSTAGES) {
              dstage = 0;
              dphase ^= 1;
            }
          }
        }
        const bool last_of_m = (t + 1 == t1) || ((t + 1) / n_vp != m);
        if (elect_one()) {
          tc_commit_pair(p_full);
          if (last_of_m) tc_commit_pair(empty_x);
        }
        __syncwarp();
      }
#ifdef FB_TIMING
      if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 76))
        printf("cta %d blend MMA warp: total %lld, wait p_empty %lld, full_d %lld (tiles %d)\n", blockIdx.x, clock64() - w_tot,
               w_pe, w_fd, t1 - t0);
#endif
    }
  } else if (warp == 3) {
    // =========================== skinning MMA issuer (leader) ===========================
    if (leader) {
      constexpr uint32_t idesc_t = umma_idesc_f16(2 * FB_VT, C::TN);
      const uint64_t desc_w0 = umma_desc_sw128(sbase + C::SM_W), desc_a0 = umma_desc_sw128(sbase + C::SM_A);
      int wbuf = 0, wphase = 0, astage = 0, aphase = 0, tb = 0, tphase = 0;
      [[maybe_unused]] long long w_te = 0, w_fa = 0, w_tot = TCLK(), tq;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(full_w + 8 * wbuf, wphase);
        const uint64_t w_desc = umma_desc_add(desc_w0, wbuf * FB_W_BYTES);
#pragma unroll 1
        for (int st = 0; st < C::NT; ++st) {
          tq = TCLK();
          mbar_wait(t_empty + 8 * tb, tphase ^ 1);
          TADD(w_te, tq);
          tq = TCLK();
          mbar_wait(full_a + 8 * astage, aphase);
          TADD(w_fa, tq);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + C::TCOL + tb * C::TN;
          const uint64_t a_desc = umma_desc_add(desc_a0, astage * C::A_HALF);
          if (elect_one()) {
            // (W window, A window) pairs of the table in k_skin_tc.cu; a window is 32 B = 2 units
            tc_mma_f16_pair(d_tmem, w_desc + 0, a_desc + 0, idesc_t, 0);
            tc_mma_f16_pair(d_tmem, w_desc + 2, a_desc + 2, idesc_t, 1);
            tc_mma_f16_pair(d_tmem, w_desc + 0, a_desc + 4, idesc_t, 1);
            tc_mma_f16_pair(d_tmem, w_desc + 4, a_desc + 0, idesc_t, 1);
            tc_mma_f16_pair(d_tmem, w_desc + 6, a_desc + 2, idesc_t, 1);
            tc_commit_pair(empty_a + 8 * astage);
            tc_commit_pair(t_full + 8 * tb);
          }
          __syncwarp();
          if (++astage == C::ASTAGES) {
            astage = 0;
            aphase ^= 1;
          }
          if (++tb == C::TBUF) {
            tb = 0;
            tphase ^= 1;
          }
        }
        if (elect_one()) tc_commit_pair(empty_w + 8 * wbuf);
        __syncwarp();
        if (++wbuf == 2) {
          wbuf = 0;
          wphase ^= 1;
        }
      }
#ifdef FB_TIMING
      if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 76))
        printf("cta %d skin MMA warp: total %lld, wait t_empty %lld, full_a %lld\n", blockIdx.x, clock64() - w_tot, w_te, w_fa);
#endif
    }
  } else if (warp >= 4) {
    // =========================== epilogue (warps 4..11, both CTAs) ===========================
    // As in k_body_tc: thread = vertex = TMEM lane; two warps per lane quarter, each 4 of a
    // skinning tile's 8 samples; the v_posed of the last PRE tiles is fetched early so that P can be
    // handed back before the super-tile ends.
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;
    constexpr int HS = C::HS;
    constexpr int PRE = C::PRE;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    const uint32_t l_p_empty = cluster_map_shared(p_empty, 0), l_t_empty = cluster_map_shared(t_empty, 0);
    int tb = 0, tphase = 0, n_tiles = 0;
    [[maybe_unused]] long long w_pf = 0, w_tf = 0, w_ld = 0, w_st = 0, w_tot = TCLK(), tq;
    for (int t = t0; t < t1; ++t, ++n_tiles) {
      const int m = t / n_vp, vt = 2 * (t % n_vp) + crank;
      const int v0 = vt * FB_VT + 32 * q;
      const bool v_ok = v0 + lane < V;
      float *const vbase = verts + ((size_t)(m * C::NS) * V + v0 + lane) * 3;
      const int b_left = B - m * C::NS;   // samples of this super-tile inside the batch

      auto load_p = [&](int s_loc, uint32_t(*pc)[HS]) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) tc_ld_32x4(lane_base + cc * C::NS + s_loc, pc[cc]);
      };
      // one skinning tile; p_in == nullptr: v_posed comes from TMEM (P), else from registers
      auto do_tile = [&](int st, const uint32_t(*p_in)[HS], bool release_p) {
        const uint32_t tcol0 = lane_base + C::TCOL + tb * C::TN + part * HS * 12;
        tq = TCLK();
        mbar_wait(t_full + 8 * tb, tphase);
        TADD(w_tf, tq);
        tc_fence_after();
        const int s_loc = st * C::ST + part * HS;   // first sample (within the super-tile) of this warp
        uint32_t r[12 * HS], pc[3][HS];
        tc_ld_32x32(tcol0, r);
        tc_ld_32x16(tcol0 + 32, r + 32);
        if (p_in == nullptr) {
          load_p(s_loc, pc);
        } else {
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int si = 0; si < HS; ++si) pc[cc][si] = p_in[cc][si];
        }
        tq = TCLK();
        tc_wait_ld();
        TADD(w_ld, tq);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(l_t_empty + 8 * tb);
          if (release_p) mbar_arrive_cluster(l_p_empty);   // this warp's last read of the super-tile's P
        }
        float o[HS][3];
#pragma unroll
        for (int si = 0; si < HS; ++si) {
          const uint32_t *T = r + 12 * si;
          const float px = __uint_as_float(pc[0][si]) * inv_scale, py = __uint_as_float(pc[1][si]) * inv_scale,
                      pz = __uint_as_float(pc[2][si]) * inv_scale;
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
            o[si][rr] = fmaf(__uint_as_float(T[4 * rr]), px,
                             fmaf(__uint_as_float(T[4 * rr + 1]), py,
                                  fmaf(__uint_as_float(T[4 * rr + 2]), pz, __uint_as_float(T[4 * rr + 3]))));
        }
        tq = TCLK();
        store_rows4(vbase + s_loc * (V * 3), V * 3, b_left - s_loc, v_ok, o[0][0], o[0][1], o[0][2], o[1][0], o[1][1], o[1][2],
                    o[2][0], o[2][1], o[2][2], o[3][0], o[3][1], o[3][2]);
        TADD(w_st, tq);
        if (++tb == C::TBUF) {
          tb = 0;
          tphase ^= 1;
        }
      };

      tq = TCLK();
      mbar_wait(p_full, n_tiles & 1);
      TADD(w_pf, tq);
      tc_fence_after();
#pragma unroll 1
      for (int st = 0; st < C::NT - PRE; ++st) do_tile(st, nullptr, PRE == 0 && st == C::NT - 1);
      if (PRE > 0) {
        uint32_t pre[PRE > 0 ? PRE : 1][3][HS];
#pragma unroll
        for (int i = 0; i < PRE; ++i) load_p((C::NT - PRE + i) * C::ST + part * HS, pre[i]);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(l_p_empty);
#pragma unroll
        for (int i = 0; i < PRE; ++i) do_tile(C::NT - PRE + i, pre[i], false);
      }
    }
#ifdef FB_TIMING
    if (lane == 0 && (warp == 4 || warp == 9) && (blockIdx.x == 0 || blockIdx.x == 1 || blockIdx.x == 76))
      printf("cta %d epilogue warp %d: total %lld, wait p_full %lld, t_full %lld, tmem ld %lld, stores %lld\n", blockIdx.x, warp,
             clock64() - w_tot, w_pf, w_tf, w_ld, w_st);
#endif
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // no CTA leaves (or frees tensor memory) while the pair's MMAs or arrivals may still touch it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
// Ring depths make no difference to the kernel alone (6 or 8 Dt16 stages, 4 A16 stages: 106-109 us).
// The early-fetch depth does, through the register count: PRE = 4 is the fastest alone (126
// registers, 106 us), PRE = 2 needs 96 registers (109 us) and leaves the kernels of the other
// contexts in flight room for twice as many co-resident warps -- the step is 2.5 % faster with it;
// PRE = 0 (78 registers, 115 us) loses more than it frees.  136 KB of shared memory.
using PairA = PairCfg<96, 8, 4, 3, 2>;

int body_pair_init(smplb_ctx *c) {
  CUDA_TRY(cudaFuncSetAttribute(k_body_pair<PairA>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairA::SM_TOTAL));
  return 0;
}

template <class C>
static int launch_pair_cfg(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts) {
  alignas(64) CUtensorMap map_x, map_a;
  TRY(tc_make_map(&map_x, 0, x16, 256, (uint64_t)B, 512, 64, C::NS / 2));
  TRY(tc_make_map(&map_a, 0, A16, 64, (uint64_t)B * 12, 128, 64, C::TN / 2));
  const int n_vp = c->Vp / (2 * FB_VT), n_m = cdiv(B, C::NS);
  const int total = n_vp * n_m;
  const int max_pairs = c->body_pairs > 0 && c->body_pairs < c->num_sms / 2 ? c->body_pairs : c->num_sms / 2;
  const int grid = 2 * (total < max_pairs ? total : max_pairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SM_TOTAL;
  cfg.stream = c->cur;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const CUtensorMap map_d = *(const CUtensorMap *)c->map_d, map_w = *(const CUtensorMap *)c->map_w;
  int Vv = c->V, Vp = c->Vp;
  float inv = c->tc_inv_scale;
  {
    ProfScope ps(c, "body_fwd_tc");
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_body_pair<C>, map_x, map_d, map_w, map_a, B, Vv, Vp, n_vp, n_m, inv, verts);
    if (e != cudaSuccess) {
      smplb_set_error("launch body_fwd_tc (CTA pairs) failed: %s", cudaGetErrorString(e));
      return SMPLB_ECUDA;
    }
  }
  c->launches++;
  return 0;
}

// verts [B][V][3]; needs an even number of 128-vertex tiles (the caller checks).
int launch_body_fwd_pair(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts) {
  return launch_pair_cfg<PairA>(c, B, x16, A16, verts);
}
