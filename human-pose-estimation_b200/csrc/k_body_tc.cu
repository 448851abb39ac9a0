// Fused blend shapes + linear blend skinning: verts straight from (x16, A16), v_posed never
// leaves the SM (batch_smpl.py:110-112, :126-132, :139-149).
//
// The two-kernel path writes v_posed (82,680 B per mesh) to HBM in k_blend_tc and reads it back in
// k_skin_tc; together they move 3x the bytes of the only tensor the caller wants (verts).  Here
// both contractions of a (128 vertices x NS samples) super-tile accumulate in TMEM with the SAME
// lane = vertex layout, so the epilogue thread of a vertex finds its v_posed and its 3x4
// transform T side by side and applies one to the other in registers:
//
//   P[v, c*NS + s]   = sum_k Dt16[c*Vp + v, k] * x16[s, k]          (blend, 2^s-scaled; K = 240)
//   T[v, 12*s' + e]  = sum_j W16[v, j] * A16[(s', e), j]            (skinning transforms, ST samples)
//   verts[s, v, r]   = T[4r..4r+2] . (2^-s P[v, :, s]) + T[4r+3]
//
// MMA shapes: blend M = 128 (vertices; "A" = a Dt16 k-block of one coordinate plane streamed through
// a ring), N = NS (samples; "B" = the x16 tile, resident for all vertex tiles of the sample block),
// 15 K-steps per plane; skinning M = 128, N = 12 * ST, five K-steps (window layout of
// k_skin_tc.cu).  An MMA of these shapes takes 56 clk when it is issued from warp-uniform code (40
// clk is the floor for any shape, tools/micro/tmem_rate.cu); the shapes are what fits the 512-column
// TMEM budget, 3 * NS (P) + TBUF * 12 * ST (T) <= 512, with two T stages.
//
// What bounds it (tools/fused_timing.py builds with -DFB_TIMING / -DFB_ABLATE, profiles/r01): the
// bytes an SM exchanges with L2.  Per launch every SM reads 5.7 MB of operands (per super-tile 192 KB
// of Dt16, 144 KB of A16, 16 KB of W16, and the x16 tile per sample block) and writes 2.3 MB of
// verts (+20 % for partially written sectors): 8.5 MB in ~177 k clk = 48 B/clk/SM.  The same rate
// shows in every ablation -- without stores 111 k clk for the reads alone; stores redirected to
// lines that stay in L2 (no HBM traffic) change nothing; neither does the way the stores are
// issued (8-byte stores, bulk copies from shared memory) nor multicasting the Dt16 stream to a CTA
// pair (the bytes still enter each SM) nor keeping W16 in TMEM (it removes shared-memory reads,
// not L2 reads).  What did help was instruction-level: warp-uniform MMA issue (157 -> 135 us),
// separate issuer warps (-> 134), the early v_posed fetch (-> 127), the stores as a real call and
// the warp index through a shuffle, both of which keep the hot loops in uniform registers
// (-> 119).  Fewer bytes per SM is what k_body_pair.cu does: cta_group::2 MMAs share the "B"
// operands (A16, x16) of a CTA pair through the pair's shared memory (-20 % of the reads, 107 us).
// That kernel is the default; this one is its single-CTA form (variant 4, and the fallback when the
// number of 128-vertex tiles is odd).
//
// Persistent, warp-specialised: warp 0 TMA producer of the blend operands, warp 1 blend MMA issuer
// and TMEM allocator, warp 2 TMA producer of the skinning operands, warp 3 skinning MMA issuer,
// warps 4-11 epilogue (two warps per TMEM lane quarter, each half of a tile's samples).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>

#include "smplb_internal.h"
#include "tc_ptx.cuh"

#ifdef FB_TIMING
#define TCLK() clock64()
#define TADD(acc, t) acc += clock64() - (t)
#else
#define TCLK() 0ll
#define TADD(acc, t)
#endif
#ifndef FB_ABLATE
#define FB_ABLATE 0   // tuning builds: 1 = epilogue only hand-shakes (MMA-side time), 2 = no MMAs (epilogue-side time),
                      // 3 = no stores, 4 = stores to lines that stay in L2, 5 = a warp's stores cover 128 contiguous bytes
                      // (wrong element order, body_common.cuh)
#endif
#include "body_common.cuh"

// NS samples per super-tile (blend MMA N), ST samples per skinning MMA, TBUF T accumulator stages.
template <int NS_, int ST_, int TBUF_, int DSTAGES_ = 4, int ASTAGES_ = 3, int PRE_ = 0, int EW_ = 2, int CL_ = 1>
struct BodyCfg {
  static constexpr int NS = NS_, ST = ST_, TBUF = TBUF_, DSTAGES = DSTAGES_, ASTAGES = ASTAGES_, PRE = PRE_, EW = EW_;
  // CL = 2: the kernel runs as clusters of two CTAs that walk the same vertex tiles on adjacent
  // sample blocks, so every Dt16 k-block is read from L2 once and multicast to both (measured: no
  // gain, the bytes still enter each SM)
  static constexpr int CL = CL_;
  static constexpr int TN = 12 * ST;            // skinning MMA N
  static constexpr int NT = NS / ST;            // skinning tiles per super-tile
  static constexpr int X_KB_BYTES = NS * 128;   // one k-block of the x16 tile
  static constexpr int X_BYTES = 4 * X_KB_BYTES;
  static constexpr int A_BYTES = TN * 128;      // A16 rows of ST samples
  static constexpr int SM_X = 0;
  static constexpr int SM_D = SM_X + X_BYTES;
  static constexpr int SM_W = SM_D + DSTAGES * FB_D_BYTES;
  static constexpr int SM_A = SM_W + 2 * FB_W_BYTES;
  static constexpr int SM_BAR = SM_A + ASTAGES * A_BYTES;
  static constexpr int SM_TOTAL = SM_BAR + 256;
  static constexpr int TCOL = 3 * NS;           // first TMEM column of the T stages
  // warps: 0 blend-operand producer, 1 blend MMA issuer, 2 skinning-operand producer, 3 skinning MMA issuer,
  // 4.. epilogue, EW warps per TMEM lane quarter, each HS = ST / EW samples of a tile.
  // PRE = skinning tiles whose v_posed the epilogue fetches early (see there)
  static constexpr int HS = ST / EW;
  static constexpr int THREADS = 32 * (4 + 4 * EW);
  static_assert(NT % TBUF == 0 && PRE < NT, "tile counts");
  static_assert(HS == 2 || HS == 4 || HS == 8, "samples per epilogue warp");
  static_assert(NS % 16 == 0 && NS % ST == 0 && ST % 4 == 0, "tile shape");
  static_assert(3 * NS + TBUF * TN <= 512, "TMEM budget");
  static_assert(X_KB_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "swizzle atoms need 1024 B alignment");
  static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
  static_assert(DSTAGES <= 4 && ASTAGES <= 4 && TBUF <= 2, "barrier slots");
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
    k_body_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_d,
              const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_a, int B, int V, int Vp,
              int n_vt, int n_m, float inv_scale, float *__restrict__ verts) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + C::SM_BAR;
  const uint32_t full_x = bar0 + 0, empty_x = bar0 + 8, p_full = bar0 + 16, p_empty = bar0 + 24;
  const uint32_t full_d = bar0 + 32, empty_d = bar0 + 64;     // C::DSTAGES (<= 4) each
  const uint32_t full_w = bar0 + 96, empty_w = bar0 + 112;    // 2 each
  const uint32_t full_a = bar0 + 128, empty_a = bar0 + 160;   // C::ASTAGES (<= 4) each
  const uint32_t t_full = bar0 + 192, t_empty = bar0 + 208;   // C::TBUF (<= 2) each
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + C::SM_BAR + 224);

  // (the shuffle tells the compiler the warp index is warp-uniform: role branches and the addresses
  // derived from it stay in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // Super-tile t of this CTA (or, CL = 2, of this cluster) -> sample block t / n_vt (times two plus
  // the CTA's rank in the pair), vertex tile t % n_vt.
  const int crank = C::CL == 2 ? (int)cluster_ctarank() : 0;
  const int n_units = gridDim.x / C::CL, unit = blockIdx.x / C::CL;
  const int total = n_vt * ((n_m + C::CL - 1) / C::CL);
  const int t0 = (int)(((long long)unit * total) / n_units);
  const int t1 = (int)(((long long)(unit + 1) * total) / n_units);

  if (threadIdx.x == 0) {
    mbar_init(full_x, 1);
    mbar_init(empty_x, 1);
    mbar_init(p_full, 1);
    mbar_init(p_empty, 4 * C::EW);         // one arrival per epilogue warp
    for (int i = 0; i < C::DSTAGES; ++i) {
      mbar_init(full_d + 8 * i, 1);
      mbar_init(empty_d + 8 * i, C::CL);   // a Dt16 stage is free when the MMAs of every CTA of the pair have read it
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
      mbar_init(t_empty + 8 * i, 4 * C::EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C::SM_BAR + 224), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (C::CL == 2) cluster_sync_all();     // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =========================== blend operand producer ===========================
    // Super-tiles are ordered sample-block major, so the x16 tile is loaded once per sample
    // block; the Dt16 k-blocks of the vertex tile (3 planes x 4 k-blocks, L2-resident) stream
    // through the ring.
    if (lane == 0) {
      int cur_m = -1, x_loads = 0, stage = 0, phase = 0, n_loads = 0;
      for (int t = t0; t < t1; ++t) {
        const int m = (t / n_vt) * C::CL + crank, vt = t % n_vt;
        if (m != cur_m) {
          if (x_loads > 0) mbar_wait(empty_x, (x_loads - 1) & 1);
          mbar_expect_tx(full_x, C::X_BYTES);
          for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + C::SM_X + kb * C::X_KB_BYTES, &map_x, kb * 64, m * C::NS, full_x);
          ++x_loads;
          cur_m = m;
        }
        for (int cc = 0; cc < 3; ++cc)
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(empty_d + 8 * stage, phase ^ 1);
            mbar_expect_tx(full_d + 8 * stage, FB_D_BYTES);
            if (C::CL == 1) {
              tma_load_2d(sbase + C::SM_D + stage * FB_D_BYTES, &map_d, kb * 64, cc * Vp + vt * FB_VT, full_d + 8 * stage);
            } else if ((n_loads & 1) == crank) {
              // the two CTAs take turns issuing; the data and the complete_tx reach both
              tma_load_2d_mc(sbase + C::SM_D + stage * FB_D_BYTES, &map_d, kb * 64, cc * Vp + vt * FB_VT, full_d + 8 * stage,
                             (uint16_t)3);
            }
            ++n_loads;
            if (++stage == C::DSTAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
      }
    }
  } else if (warp == 2) {
    // =========================== skinning operand producer ===========================
    if (lane == 0) {
      int wbuf = 0, wphase = 0, stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int m = (t / n_vt) * C::CL + crank, vt = t % n_vt;
        mbar_wait(empty_w + 8 * wbuf, wphase ^ 1);
        mbar_expect_tx(full_w + 8 * wbuf, FB_W_BYTES);
        tma_load_2d(sbase + C::SM_W + wbuf * FB_W_BYTES, &map_w, 0, vt * FB_VT, full_w + 8 * wbuf);
        if (++wbuf == 2) {
          wbuf = 0;
          wphase ^= 1;
        }
        for (int st = 0; st < C::NT; ++st) {
          mbar_wait(empty_a + 8 * stage, phase ^ 1);
          mbar_expect_tx(full_a + 8 * stage, C::A_BYTES);
          tma_load_2d(sbase + C::SM_A + stage * C::A_BYTES, &map_a, 0, (m * C::NS + st * C::ST) * 12, full_a + 8 * stage);
          if (++stage == C::ASTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== blend MMA issuer ===========================
    // All 32 lanes run the loop and wait on the barriers; the elected lane issues (see elect_one).
    // The two contractions have their own issuing warps (1: blend, 3: skinning): a warp's loop
    // bookkeeping (~200 clk per group of MMAs) does not overlap its own issue stalls, so one warp
    // doing both kept the tensor pipe idle half of the time.
    {
      constexpr uint32_t idesc_p = umma_idesc_f16(FB_VT, C::NS);
      const uint64_t desc_d0 = umma_desc_sw128(sbase + C::SM_D), desc_x0 = umma_desc_sw128(sbase + C::SM_X);
      int cur_m = -1, x_loads = 0, dstage = 0, dphase = 0, n_tiles = 0;
      [[maybe_unused]] long long w_pe = 0, w_fd = 0, w_tot = TCLK(), tq;
      for (int t = t0; t < t1; ++t, ++n_tiles) {
        const int m = t / n_vt;                       // (the pair's sample-block pair)
        if (m != cur_m) {
          mbar_wait(full_x, x_loads & 1);
          ++x_loads;
          cur_m = m;
        }
        // P = Dt16 tile x x16 tile^T, three coordinate planes
        tq = TCLK();
        mbar_wait(p_empty, (n_tiles & 1) ^ 1);       // the epilogue has read the previous P
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
            const uint64_t b_desc = umma_desc_add(desc_x0, kb * C::X_KB_BYTES);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kb == 3 && k == 3) continue;      // K = 240: the last 16 columns are zero padding
                if (FB_ABLATE != 2) tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_p, (kb | k) != 0);
              }
              if (C::CL == 1) tc_commit(empty_d + 8 * dstage);
              else tc_commit_mc(empty_d + 8 * dstage, (uint16_t)3);   // releases the stage in both CTAs
            }
            __syncwarp();
            if (++dstage == C::DSTAGES) {
              dstage = 0;
              dphase ^= 1;
            }
          }
        }
        const bool last_of_m = (t + 1 == t1) || ((t + 1) / n_vt != m);
        if (elect_one()) {
          tc_commit(p_full);
          if (last_of_m) tc_commit(empty_x);
        }
        __syncwarp();
      }
#ifdef FB_TIMING
      if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
        printf("cta %d blend MMA warp: total %lld, wait p_empty %lld, full_d %lld (tiles %d)\n", blockIdx.x, clock64() - w_tot,
               w_pe, w_fd, t1 - t0);
#endif
    }
  } else if (warp == 3) {
    // =========================== skinning MMA issuer ===========================
    {
      constexpr uint32_t idesc_t = umma_idesc_f16(FB_VT, C::TN);
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
          const uint64_t a_desc = umma_desc_add(desc_a0, astage * C::A_BYTES);
          if (elect_one()) {
            // (W window, A window) pairs of the table in k_skin_tc.cu; a window is 32 B = 2 units
            if (FB_ABLATE != 2) {
              tc_mma_f16(d_tmem, w_desc + 0, a_desc + 0, idesc_t, 0);
              tc_mma_f16(d_tmem, w_desc + 2, a_desc + 2, idesc_t, 1);
              tc_mma_f16(d_tmem, w_desc + 0, a_desc + 4, idesc_t, 1);
              tc_mma_f16(d_tmem, w_desc + 4, a_desc + 0, idesc_t, 1);
              tc_mma_f16(d_tmem, w_desc + 6, a_desc + 2, idesc_t, 1);
            }
            tc_commit(empty_a + 8 * astage);
            tc_commit(t_full + 8 * tb);
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
        if (elect_one()) tc_commit(empty_w + 8 * wbuf);
        __syncwarp();
        if (++wbuf == 2) {
          wbuf = 0;
          wphase ^= 1;
        }
      }
#ifdef FB_TIMING
      if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
        printf("cta %d skin MMA warp: total %lld, wait t_empty %lld, full_a %lld\n", blockIdx.x, clock64() - w_tot, w_te, w_fa);
#endif
    }
  } else if (warp >= 4) {
    // =========================== epilogue (warps 4..) ===========================
    // EW warps per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31), each taking
    // HS = ST / EW samples of a skinning tile.  Thread = vertex: T (12 columns per sample) and the
    // three coordinates of v_posed come out of TMEM; nothing passes through shared memory.
    // P has a single TMEM stage, so the blend of the next super-tile cannot start before the
    // last read of this one: the v_posed values of the last PRE tiles are therefore fetched into
    // registers early and P is handed back PRE tiles before the super-tile ends, which hides most
    // of the blend behind the remaining skinning tiles.
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;             // which HS samples of a tile
    constexpr int HS = C::HS;
    constexpr int PRE = C::PRE;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    int tb = 0, tphase = 0, n_tiles = 0;
    [[maybe_unused]] long long w_pf = 0, w_tf = 0, w_ld = 0, w_tot = TCLK(), tq;
    for (int t = t0; t < t1; ++t, ++n_tiles) {
      const int m = (t / n_vt) * C::CL + crank, vt = t % n_vt;
      const int v0 = vt * FB_VT + 32 * q;
      const bool v_ok = v0 + lane < V;
      // (FB_ABLATE == 4: every super-tile writes the rows of sample block 0 -- same store instructions and
      // L2 traffic, but the lines are overwritten in L2 and never reach HBM)
      float *const vbase = verts + ((size_t)((FB_ABLATE == 4 ? 0 : m) * C::NS) * V + v0 + lane) * 3;
      const int b_left = FB_ABLATE == 4 ? C::NS : B - m * C::NS;   // samples of this super-tile inside the batch (<= 0: a padding block)

      // the three coordinates of HS samples' v_posed (columns s_loc.. of the three planes of P)
      auto load_p = [&](int s_loc, uint32_t(*pc)[HS]) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          const uint32_t col = lane_base + cc * C::NS + s_loc;
          if (HS == 2) tc_ld_32x2(col, pc[cc]);
          if (HS == 4) tc_ld_32x4(col, pc[cc]);
          if (HS == 8) tc_ld_32x8(col, pc[cc]);
        }
      };
      // one skinning tile; p_in == nullptr: v_posed comes from TMEM (P), else from registers
      auto do_tile = [&](int st, const uint32_t(*p_in)[HS], bool release_p) {
        const uint32_t my_full = t_full + 8 * tb, my_empty = t_empty + 8 * tb;
        const uint32_t tcol0 = lane_base + C::TCOL + tb * C::TN + part * HS * 12;
        tq = TCLK();
        mbar_wait(my_full, tphase);
        TADD(w_tf, tq);
        tc_fence_after();
        const int s_loc = st * C::ST + part * HS;   // first sample (within the super-tile) of this warp
        uint32_t r[12 * HS], pc[3][HS];
        if (HS == 2) {
          tc_ld_32x16(tcol0, r);
          tc_ld_32x8(tcol0 + 16, r + 16);
        }
        if (HS == 4) {
          tc_ld_32x32(tcol0, r);
          tc_ld_32x16(tcol0 + 32, r + 32);
        }
        if (HS == 8) {
          tc_ld_32x32(tcol0, r);
          tc_ld_32x32(tcol0 + 32, r + 32);
          tc_ld_32x32(tcol0 + 64, r + 64);
        }
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
          mbar_arrive(my_empty);
          if (release_p) mbar_arrive(p_empty);   // this warp's last read of the super-tile's P
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
        if (FB_ABLATE == 3) {
        } else {
          // verts[b][v][xyz]: three scalar stores 12 B apart per sample.  A build without the stores
          // is 37 % faster, but how they are issued is not the lever: 8-byte stores (transposed through
          // shared memory +28 %, lane pairs with one shuffle +7 %) and rows staged in shared memory and
          // written with cp.async.bulk copies (+17 %; rows are 8-byte aligned, so odd rows need a
          // scalar head and tail) were all measured slower than this
          if (HS == 4) {
            store_rows4(vbase + s_loc * (V * 3), V * 3, b_left - s_loc, v_ok, o[0][0], o[0][1], o[0][2], o[1][0], o[1][1], o[1][2],
                        o[2 % HS][0], o[2 % HS][1], o[2 % HS][2], o[3 % HS][0], o[3 % HS][1], o[3 % HS][2]);
          } else {
#pragma unroll
            for (int si = 0; si < HS; ++si) {
              const int sl = s_loc + si;
              if (sl < b_left && v_ok) {
                float *dst = vbase + sl * (V * 3);
                FB_ST(dst, o[si][0]);
                FB_ST(dst + 1, o[si][1]);
                FB_ST(dst + 2, o[si][2]);
              }
            }
          }
        }
        if (++tb == C::TBUF) {
          tb = 0;
          tphase ^= 1;
        }
      };

      tq = TCLK();
      mbar_wait(p_full, n_tiles & 1);
      TADD(w_pf, tq);
      tc_fence_after();
      if (FB_ABLATE == 1) {
        for (int st = 0; st < C::NT; ++st) {
          mbar_wait(t_full + 8 * tb, tphase);
          if (lane == 0) {
            mbar_arrive(t_empty + 8 * tb);
            if (st == C::NT - 1) mbar_arrive(p_empty);
          }
          if (++tb == C::TBUF) {
            tb = 0;
            tphase ^= 1;
          }
        }
        continue;
      }
#pragma unroll 1
      for (int st = 0; st < C::NT - PRE; ++st) do_tile(st, nullptr, PRE == 0 && st == C::NT - 1);
      if (PRE > 0) {
        uint32_t pre[PRE > 0 ? PRE : 1][3][HS];
#pragma unroll
        for (int i = 0; i < PRE; ++i) load_p((C::NT - PRE + i) * C::ST + part * HS, pre[i]);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_empty);
#pragma unroll
        for (int i = 0; i < PRE; ++i) do_tile(C::NT - PRE + i, pre[i], false);
      }
    }
#ifdef FB_TIMING
    if (lane == 0 && (warp == 4 || warp == 9) && (blockIdx.x == 0 || blockIdx.x == 77))
      printf("cta %d epilogue warp %d: total %lld, wait p_full %lld, t_full %lld, tmem ld %lld\n", blockIdx.x, warp,
             clock64() - w_tot, w_pf, w_tf, w_ld);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (C::CL == 2) cluster_sync_all();     // no CTA leaves while its peer may still multicast to it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// Variant with the W16 tile as a TMEM-resident "A" operand (tcgen05.mma with A in tensor memory);
// smplb_debug_set("fused", 5).  Measured equal to k_body_tc (127.0 vs 127.8 us), so it is not the
// default; it stays as the tested reference for TS-mode MMAs and tcgen05.st in this code base.
// In k_body_tc the skinning MMAs re-read the 16 KB W16 tile from shared memory for every one of
// their 60 K-steps per super-tile (240 KB of the ~1.1 MB of shared-memory traffic that bounds the
// kernel).  A W16 row is 64 fp16 = 32 TMEM columns -- exactly what 3 * 96 + 2 * 96 leaves free -- so
// the epilogue threads (thread = vertex = TMEM lane) copy their row there with tcgen05.st whenever
// the vertex tile changes.  To make that rare the super-tiles are ordered vertex-tile major (a CTA
// walks ~16 consecutive sample blocks of one vertex tile) and the x16 tile, which now changes every
// super-tile, is double-buffered in the space the W16 ring used.
struct BodyW {
  static constexpr int NS = 96, ST = 8, TBUF = 2, DSTAGES = 4, ASTAGES = 3, PRE = 4;
  static constexpr int TN = 12 * ST, NT = NS / ST;
  static constexpr int X_KB_BYTES = NS * 128, X_BYTES = 4 * X_KB_BYTES;
  static constexpr int A_BYTES = TN * 128;
  static constexpr int SM_X = 0;                                  // two x16 tiles
  static constexpr int SM_D = SM_X + 2 * X_BYTES;
  static constexpr int SM_A = SM_D + DSTAGES * FB_D_BYTES;
  static constexpr int SM_BAR = SM_A + ASTAGES * A_BYTES;
  static constexpr int SM_TOTAL = SM_BAR + 256;
  static constexpr int TCOL = 3 * NS;                             // T stages
  static constexpr int WCOL = TCOL + TBUF * TN;                   // W16 tile: 32 columns
  static constexpr int THREADS = 32 * 12;
  static_assert(WCOL + 32 <= 512, "TMEM budget");
  static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
};

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns from registers: thread i writes row (lane base + i)
__device__ __forceinline__ void tc_st_32x16(uint32_t taddr, const uint32_t *r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__global__ void __launch_bounds__(BodyW::THREADS, 1)
    k_body_wt(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_d,
              const __grid_constant__ CUtensorMap map_a, const uint4 *__restrict__ W16, int B, int V, int Vp, int n_vt,
              int n_m, float inv_scale, float *__restrict__ verts) {
  using C = BodyW;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + C::SM_BAR;
  const uint32_t full_x = bar0 + 0, empty_x = bar0 + 16;      // 2 each
  const uint32_t full_d = bar0 + 32, empty_d = bar0 + 64;     // DSTAGES (<= 4) each
  const uint32_t p_full = bar0 + 96, p_empty = bar0 + 104, w_ready = bar0 + 112;
  const uint32_t full_a = bar0 + 128, empty_a = bar0 + 160;   // ASTAGES (<= 4) each
  const uint32_t t_full = bar0 + 192, t_empty = bar0 + 208;   // TBUF (<= 2) each
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + C::SM_BAR + 224);

  // (the shuffle tells the compiler the warp index is warp-uniform: role branches and the addresses
  // derived from it stay in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int total = n_vt * n_m;
  const int t0 = (int)(((long long)blockIdx.x * total) / gridDim.x);
  const int t1 = (int)(((long long)(blockIdx.x + 1) * total) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(full_x + 8 * i, 1);
      mbar_init(empty_x + 8 * i, 1);
    }
    mbar_init(p_full, 1);
    mbar_init(p_empty, 8);                 // one arrival per epilogue warp
    mbar_init(w_ready, 8);
    for (int i = 0; i < C::DSTAGES; ++i) {
      mbar_init(full_d + 8 * i, 1);
      mbar_init(empty_d + 8 * i, 1);
    }
    for (int i = 0; i < C::ASTAGES; ++i) {
      mbar_init(full_a + 8 * i, 1);
      mbar_init(empty_a + 8 * i, 1);
    }
    for (int i = 0; i < C::TBUF; ++i) {
      mbar_init(t_full + 8 * i, 1);
      mbar_init(t_empty + 8 * i, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + C::SM_BAR + 224), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // super-tile t -> vertex tile t / n_m, sample block t % n_m
  if (warp == 0) {
    // =========================== blend operand producer ===========================
    if (lane == 0) {
      int stage = 0, phase = 0, n_tiles = 0;
      for (int t = t0; t < t1; ++t, ++n_tiles) {
        const int vt = t / n_m, m = t % n_m;
        const int xb = n_tiles & 1;
        mbar_wait(empty_x + 8 * xb, ((n_tiles >> 1) & 1) ^ 1);
        mbar_expect_tx(full_x + 8 * xb, C::X_BYTES);
        for (int kb = 0; kb < 4; ++kb)
          tma_load_2d(sbase + C::SM_X + xb * C::X_BYTES + kb * C::X_KB_BYTES, &map_x, kb * 64, m * C::NS, full_x + 8 * xb);
        for (int cc = 0; cc < 3; ++cc)
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(empty_d + 8 * stage, phase ^ 1);
            mbar_expect_tx(full_d + 8 * stage, FB_D_BYTES);
            tma_load_2d(sbase + C::SM_D + stage * FB_D_BYTES, &map_d, kb * 64, cc * Vp + vt * FB_VT, full_d + 8 * stage);
            if (++stage == C::DSTAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
      }
    }
  } else if (warp == 2) {
    // =========================== skinning operand producer (A16 rows) ===========================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int m = t % n_m;
        for (int st = 0; st < C::NT; ++st) {
          mbar_wait(empty_a + 8 * stage, phase ^ 1);
          mbar_expect_tx(full_a + 8 * stage, C::A_BYTES);
          tma_load_2d(sbase + C::SM_A + stage * C::A_BYTES, &map_a, 0, (m * C::NS + st * C::ST) * 12, full_a + 8 * stage);
          if (++stage == C::ASTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== blend MMA issuer ===========================
    constexpr uint32_t idesc_p = umma_idesc_f16(FB_VT, C::NS);
    const uint64_t desc_d0 = umma_desc_sw128(sbase + C::SM_D), desc_x0 = umma_desc_sw128(sbase + C::SM_X);
    int dstage = 0, dphase = 0, n_tiles = 0;
    for (int t = t0; t < t1; ++t, ++n_tiles) {
      const int xb = n_tiles & 1;
      mbar_wait(full_x + 8 * xb, (n_tiles >> 1) & 1);
      mbar_wait(p_empty, (n_tiles & 1) ^ 1);       // the epilogue has read the previous P
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 3; ++cc) {
        const uint32_t d_tmem = tmem_base + cc * C::NS;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(full_d + 8 * dstage, dphase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_add(desc_d0, dstage * FB_D_BYTES);
          const uint64_t b_desc = umma_desc_add(desc_x0, xb * C::X_BYTES + kb * C::X_KB_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kb == 3 && k == 3) continue;      // K = 240: the last 16 columns are zero padding
              tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_p, (kb | k) != 0);
            }
            tc_commit(empty_d + 8 * dstage);
          }
          __syncwarp();
          if (++dstage == C::DSTAGES) {
            dstage = 0;
            dphase ^= 1;
          }
        }
      }
      if (elect_one()) {
        tc_commit(p_full);
        tc_commit(empty_x + 8 * xb);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // =========================== skinning MMA issuer (A = W16 in TMEM) ===========================
    constexpr uint32_t idesc_t = umma_idesc_f16(FB_VT, C::TN);
    const uint64_t desc_a0 = umma_desc_sw128(sbase + C::SM_A);
    const uint32_t w_tmem = tmem_base + C::WCOL;     // window w of the W16 row = columns 8 w .. 8 w + 7
    int astage = 0, aphase = 0, tb = 0, tphase = 0, cur_vt = -1, w_loads = 0;
    for (int t = t0; t < t1; ++t) {
      const int vt = t / n_m;
      if (vt != cur_vt) {
        mbar_wait(w_ready, w_loads & 1);             // the epilogue has put this vertex tile's W16 rows in TMEM
        ++w_loads;
        cur_vt = vt;
        tc_fence_after();
      }
#pragma unroll 1
      for (int st = 0; st < C::NT; ++st) {
        mbar_wait(t_empty + 8 * tb, tphase ^ 1);
        mbar_wait(full_a + 8 * astage, aphase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + C::TCOL + tb * C::TN;
        const uint64_t a_desc = umma_desc_add(desc_a0, astage * C::A_BYTES);
        if (elect_one()) {
          // (W window, A window) pairs of the table in k_skin_tc.cu
          tc_mma_f16_ts(d_tmem, w_tmem + 0, a_desc + 0, idesc_t, 0);
          tc_mma_f16_ts(d_tmem, w_tmem + 8, a_desc + 2, idesc_t, 1);
          tc_mma_f16_ts(d_tmem, w_tmem + 0, a_desc + 4, idesc_t, 1);
          tc_mma_f16_ts(d_tmem, w_tmem + 16, a_desc + 0, idesc_t, 1);
          tc_mma_f16_ts(d_tmem, w_tmem + 24, a_desc + 2, idesc_t, 1);
          tc_commit(empty_a + 8 * astage);
          tc_commit(t_full + 8 * tb);
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
    }
  } else if (warp >= 4) {
    // =========================== epilogue (warps 4..11), as in k_body_tc ===========================
    const int q = warp & 3;
    const int half = ((warp - 4) >> 2) & 1;
    constexpr int PRE = C::PRE;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    int tb = 0, tphase = 0, n_tiles = 0, cur_vt = -1;
    for (int t = t0; t < t1; ++t, ++n_tiles) {
      const int vt = t / n_m, m = t % n_m;
      const int v0 = vt * FB_VT + 32 * q;
      const bool v_ok = v0 + lane < V;
      float *const vbase = verts + ((size_t)(m * C::NS) * V + v0 + lane) * 3;
      const int b_left = B - m * C::NS;
      if (vt != cur_vt) {
        // Every skinning MMA of the previous vertex tile has completed (this warp has waited for the
        // T of its last tile): copy this thread's half of its W16 row (rows >= V are zero) to TMEM.
        cur_vt = vt;
        uint32_t wr[16];
        const uint4 *src = W16 + (size_t)(v0 + lane) * 8 + half * 4;   // 128 B per row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 x = src[i];
          wr[4 * i + 0] = x.x;
          wr[4 * i + 1] = x.y;
          wr[4 * i + 2] = x.z;
          wr[4 * i + 3] = x.w;
        }
        tc_st_32x16(lane_base + C::WCOL + half * 16, wr);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(w_ready);
      }

      auto do_tile = [&](int st, const uint32_t(*p_in)[4], bool release_p) {
        const uint32_t my_full = t_full + 8 * tb, my_empty = t_empty + 8 * tb;
        const uint32_t tcol0 = lane_base + C::TCOL + tb * C::TN + half * 48;
        mbar_wait(my_full, tphase);
        tc_fence_after();
        const int s_loc = st * C::ST + half * 4;
        uint32_t r[48], pc[3][4];
        tc_ld_32x32(tcol0, r);
        tc_ld_32x16(tcol0 + 32, r + 32);
        if (p_in == nullptr) {
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) tc_ld_32x4(lane_base + cc * C::NS + s_loc, pc[cc]);
        } else {
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int si = 0; si < 4; ++si) pc[cc][si] = p_in[cc][si];
        }
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(my_empty);
          if (release_p) mbar_arrive(p_empty);
        }
#pragma unroll
        for (int si = 0; si < 4; ++si) {
          const uint32_t *T = r + 12 * si;
          const float px = __uint_as_float(pc[0][si]) * inv_scale, py = __uint_as_float(pc[1][si]) * inv_scale,
                      pz = __uint_as_float(pc[2][si]) * inv_scale;
          float o[3];
#pragma unroll
          for (int rr = 0; rr < 3; ++rr)
            o[rr] = fmaf(__uint_as_float(T[4 * rr]), px,
                         fmaf(__uint_as_float(T[4 * rr + 1]), py,
                              fmaf(__uint_as_float(T[4 * rr + 2]), pz, __uint_as_float(T[4 * rr + 3]))));
          const int sl = s_loc + si;
          if (sl < b_left && v_ok) {
            float *dst = vbase + sl * (V * 3);
            __stcs(dst, o[0]);
            __stcs(dst + 1, o[1]);
            __stcs(dst + 2, o[2]);
          }
        }
        if (++tb == C::TBUF) {
          tb = 0;
          tphase ^= 1;
        }
      };

      mbar_wait(p_full, n_tiles & 1);
      tc_fence_after();
#pragma unroll 1
      for (int st = 0; st < C::NT - PRE; ++st) do_tile(st, nullptr, false);
      uint32_t pre[PRE][3][4];
#pragma unroll
      for (int i = 0; i < PRE; ++i)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) tc_ld_32x4(lane_base + cc * C::NS + (C::NT - PRE + i) * C::ST + half * 4, pre[i][cc]);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_empty);
#pragma unroll
      for (int i = 0; i < PRE; ++i) do_tile(C::NT - PRE + i, pre[i], false);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
// operand rows of 16-bit elements, box = 64 columns (one swizzle atom) x box_rows
static int make_map_rows16(CUtensorMap *map, const void *ptr, uint64_t row_halves, uint64_t rows, uint32_t box_rows) {
  return tc_make_map(map, 0, ptr, row_halves, rows, row_halves * 2, 64, box_rows);
}

// Needs the operands of both tensor-core kernels (Dt16 + its scale, W16).
using BodyA = BodyCfg<96, 8, 2, 4, 3, 4>;   // best single-CTA configuration: double-buffered T, v_posed of the last 4 tiles fetched early
using BodyB = BodyCfg<96, 8, 2, 4, 3, 0>;   // no early fetch (the blend is exposed)
using BodyC = BodyCfg<128, 4, 2, 4, 4, 8>;   // 128-sample super-tiles (25 % fewer Dt16 bytes per sample), two 4-sample T stages
using BodyP = BodyCfg<96, 8, 2, 4, 3, 4, 2, 2>;  // CTA pairs, Dt16 multicast

int body_tc_init(smplb_ctx *c) {
  c->body_tc_ok = false;
  if (!c->tc_ok || !c->skin_tc_ok) return 0;
  CUDA_TRY(cudaFuncSetAttribute(k_body_tc<BodyA>, cudaFuncAttributeMaxDynamicSharedMemorySize, BodyA::SM_TOTAL));
  CUDA_TRY(cudaFuncSetAttribute(k_body_tc<BodyB>, cudaFuncAttributeMaxDynamicSharedMemorySize, BodyB::SM_TOTAL));
  CUDA_TRY(cudaFuncSetAttribute(k_body_tc<BodyC>, cudaFuncAttributeMaxDynamicSharedMemorySize, BodyC::SM_TOTAL));
  CUDA_TRY(cudaFuncSetAttribute(k_body_wt, cudaFuncAttributeMaxDynamicSharedMemorySize, BodyW::SM_TOTAL));
  CUDA_TRY(cudaFuncSetAttribute(k_body_tc<BodyP>, cudaFuncAttributeMaxDynamicSharedMemorySize, BodyP::SM_TOTAL));
  TRY(body_pair_init(c));
  TRY(body_res_init(c));
  c->body_tc_ok = true;
  return 0;
}

template <class C>
static int launch_body_cfg(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts) {
  alignas(64) CUtensorMap map_x, map_a;
  TRY(make_map_rows16(&map_x, x16, 256, (uint64_t)B, C::NS));
  TRY(make_map_rows16(&map_a, A16, 64, (uint64_t)B * 12, C::TN));
  const int n_vt = c->Vp / FB_VT, n_m = cdiv(B, C::NS);
  const int total = n_vt * n_m;
  const int grid = total < c->num_sms ? total : c->num_sms;
  LAUNCH(c, "body_fwd_tc", grid, C::THREADS, C::SM_TOTAL, k_body_tc<C>, map_x, *(const CUtensorMap *)c->map_d,
         *(const CUtensorMap *)c->map_w, map_a, B, c->V, c->Vp, n_vt, n_m, c->tc_inv_scale, verts);
  return 0;
}

// CTA-pair variant: launched as clusters of two (the grid is rounded down to an even CTA count).
static int launch_body_pair(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts) {
  using C = BodyP;
  alignas(64) CUtensorMap map_x, map_a;
  TRY(make_map_rows16(&map_x, x16, 256, (uint64_t)B, C::NS));
  TRY(make_map_rows16(&map_a, A16, 64, (uint64_t)B * 12, C::TN));
  const int n_vt = c->Vp / FB_VT, n_m = cdiv(B, C::NS);
  const int total = n_vt * cdiv(n_m, 2);
  int grid = 2 * (total < c->num_sms / 2 ? total : c->num_sms / 2);
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_body_tc<C>, map_x, map_d, map_w, map_a, B, Vv, Vp, n_vt, n_m, inv, verts);
    if (e != cudaSuccess) {
      smplb_set_error("launch body_fwd_tc (cluster) failed: %s", cudaGetErrorString(e));
      return SMPLB_ECUDA;
    }
  }
  c->launches++;
  return 0;
}

static int launch_body_wt(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts) {
  alignas(64) CUtensorMap map_x, map_a;
  TRY(make_map_rows16(&map_x, x16, 256, (uint64_t)B, BodyW::NS));
  TRY(make_map_rows16(&map_a, A16, 64, (uint64_t)B * 12, BodyW::TN));
  const int n_vt = c->Vp / FB_VT, n_m = cdiv(B, BodyW::NS);
  const int total = n_vt * n_m;
  const int grid = total < c->num_sms ? total : c->num_sms;
  LAUNCH(c, "body_fwd_tc", grid, BodyW::THREADS, BodyW::SM_TOTAL, k_body_wt, map_x, *(const CUtensorMap *)c->map_d, map_a,
         (const uint4 *)c->d_W16, B, c->V, c->Vp, n_vt, n_m, c->tc_inv_scale, verts);
  return 0;
}

// verts [B][V][3] from the operand rows pose_fwd wrote (x16 [B][256], A16 [12 B][64]).
// Default: the CTA-pair kernel with the resident Dt16 tile of k_body_res.cu (it needs an even number
// of 128-vertex tiles; SMPL has 54).  smplb_debug_set("fused", 9) selects its sixteen-epilogue-warp
// configuration, 7 the streaming CTA-pair kernel of k_body_pair.cu, 2 .. 6 the single-CTA
// configurations of this file (tuning / validation; 4 = the best of them, which is also the fallback).
int launch_body_fwd_tc(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts) {
  RET_IF(!c->body_tc_ok, SMPLB_ESTATE, "fused tcgen05 blend+skinning path is not initialised");
  switch (c->use_fused) {
    case 2: return launch_body_cfg<BodyB>(c, B, x16, A16, verts);
    case 3: return launch_body_cfg<BodyC>(c, B, x16, A16, verts);
    case 4: return launch_body_cfg<BodyA>(c, B, x16, A16, verts);
    case 5: return launch_body_wt(c, B, x16, A16, verts);
    case 6: return launch_body_pair(c, B, x16, A16, verts);
    case 7:
      if ((c->Vp / FB_VT) % 2 == 0) return launch_body_fwd_pair(c, B, x16, A16, verts);
      return launch_body_cfg<BodyA>(c, B, x16, A16, verts);
    case 9:
    case 10:
    default:
      if ((c->Vp / FB_VT) % 2 == 0) return launch_body_fwd_res(c, B, x16, A16, verts, c->use_fused);
      return launch_body_cfg<BodyA>(c, B, x16, A16, verts);
  }
}
