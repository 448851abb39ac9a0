// Fused blend shapes + linear blend skinning on CTA pairs with the blend operand RESIDENT in shared
// memory (batch_smpl.py:110-112, :126-132, :139-149).
//
// k_body_pair moves its bytes between the SMs and L2 at the rate the L2 slices sustain (~10 TB/s at
// the clocks it runs at), and 455 of the 671 MB it reads per launch at B = 4096 are the Dt16 tile of
// its vertices, streamed again for every 96-sample block.  That tile is 3 planes x 128 vertices x
// 256 fp16 = 192 KB -- it fits in one SM's shared memory if everything else gets out of the way:
//
//   * super-tiles are walked vertex-tile major: a CTA pair takes ~16 consecutive sample blocks of ONE
//     pair of vertex tiles, so the Dt16 tile is loaded once (twice for a pair whose range straddles
//     two vertex tiles) instead of 16 times;
//   * the W16 tile (skinning weights, the "A" operand of T = W.A) lives in tensor memory: a row is 64
//     fp16 = 32 columns, exactly what P (3 x 96) and the two T stages (2 x 96) leave of the 512; the
//     epilogue threads (thread = vertex = TMEM lane) copy their row there with tcgen05.st when the
//     vertex tile changes;
//   * the "B" operands of both contractions -- the four k-blocks of the x16 tile and the twelve A16
//     tiles of a super-tile, all [48 rows x 128 B] halves of a cta_group::2 operand -- stream through
//     two small rings of 6 KB stages (x16: 3, A16: 4); the blend MMAs run k-block major so that an
//     x16 stage is free again after its three planes.  (One ring shared by the two MMA issuers does
//     not work: an issuer that skips the other's items can get two ring revolutions away from the
//     producer, where an mbarrier's phase parity is ambiguous -- measured as wrong vertices.)
//
//   * K = 240 (207 pose + 30 split shape + 3 split template columns), so the last 64-column k-block is only 3/4 used:
//     it is loaded as three [128 rows x 16 columns] SWIZZLE_32B slabs (one K = 16 step each) instead of one
//     SWIZZLE_128B block, which frees 12 KB for two more ring stages (x16: 3, A16: 4): the latency of the ring loads
//     under the kernel's own store traffic is what the hand-shake chain waits for (112.6 -> 105.5 us on one box; verts
//     stores through staging buffers and the TMA engine with those 12 KB instead were measured equal to st.global,
//     profiles/r02/vertex_kernel_ablations.md).
//
// Shared memory: 180 KB + 42 KB of rings + barriers.  Per launch at B = 4096 a CTA reads ~1.9 MB instead of
// 4.5 MB from L2.  Roles per CTA: warp 0 ring producer, warp 2 Dt16-tile producer (both CTAs load
// their own halves; complete_tx goes to the leader's barriers), warps 1 / 3 blend / skinning MMA
// issuers (leader only, commits multicast to both CTAs), warps 4-11 epilogue.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>

#include "smplb_internal.h"
#include "tc_ptx.cuh"
#include "body_common.cuh"

// Tuning builds of this kernel (tools/build_variant.sh <name> -DFB_ABL2=<bits>): 1 no MMAs, 2 no TMEM loads, 4 no stores,
// 8 stores go to the rows of sample block 0 (they stay in L2), 16 no TMA loads into the x16 / A16 rings.
#ifndef FB_ABL2
#define FB_ABL2 0
#endif
#ifdef FB_TIMING
#define TCLK() clock64()
#define TADD(acc, t) acc += clock64() - (t)
#else
#define TCLK() 0ll
#define TADD(acc, t)
#endif

// D[tmem] (+)= A[tmem] * B[smem] over a CTA pair: each CTA's 128 rows of A come from its own tensor memory
__device__ __forceinline__ void tc_mma_f16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns from registers: thread i writes row (lane base + i)
__device__ __forceinline__ void tc_st_32x16_r(uint32_t taddr, const uint32_t *r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// K-major SWIZZLE_32B operand ([rows x 32 B], 8-row groups 256 B apart): one K = 16 step per descriptor
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}

// EG = epilogue warp groups: 1 = eight warps walk every skinning tile; 2 = sixteen warps, group g takes the
// tiles (and the T stage) of parity g, so one group's stores overlap the other's TMEM loads and FMAs.
// PRE = tiles per group whose v_posed is fetched early (P is handed back before the super-tile ends).
template <int NS_, int ST_, int XSTAGES_, int ASTAGES_, int PRE_, int EG_>
struct ResCfg {
  static constexpr int NS = NS_, ST = ST_, TBUF = 2, XSTAGES = XSTAGES_, ASTAGES = ASTAGES_, PRE = PRE_, EW = 2, EG = EG_;
  static constexpr int TN = 12 * ST;                 // skinning MMA N
  static constexpr int NT = NS / ST;                 // skinning tiles per super-tile
  static constexpr int X_BYTES = (NS / 2) * 128;     // one stage of the x16 ring: this CTA's rows of one k-block
  static constexpr int A_BYTES = (TN / 2) * 128;     // one stage of the A16 ring: this CTA's rows of one skinning tile
  static constexpr int D_MAIN = 9 * FB_D_BYTES;      // 3 planes x k-blocks 0..2 of this CTA's vertex tile, SWIZZLE_128B
  static constexpr int D_SLAB = FB_VT * 32;          // one K = 16 step of the last k-block: [128 rows x 32 B], SWIZZLE_32B
  static constexpr int D_TILE = D_MAIN + 9 * D_SLAB; // 180 KB: the zero padding of K = 240 to 256 never enters shared memory
  static constexpr int SM_D = 0;
  static constexpr int SM_X = SM_D + D_TILE;
  static constexpr int SM_A = SM_X + XSTAGES * X_BYTES;
  static constexpr int SM_BAR = SM_A + ASTAGES * A_BYTES;
  static constexpr int SM_TOTAL = SM_BAR + 512;
  static constexpr int TCOL = 3 * NS;                // first TMEM column of the T stages
  static constexpr int WCOL = TCOL + TBUF * TN;      // W16 tile: 32 columns
  static constexpr int HS = ST / EW;                 // samples per epilogue warp and tile
  static constexpr int THREADS = 32 * (4 + 4 * EW * EG);
  static constexpr int GT = NT / EG;                 // tiles per group and super-tile
  // shared-memory offset of (plane cc, k-block kb) of the resident tile; kb = 3: the first of the plane's three slabs
  __host__ __device__ static constexpr int d_off(int cc, int kb) { return kb < 3 ? (cc * 3 + kb) * FB_D_BYTES : D_MAIN + cc * 3 * D_SLAB; }
  static_assert(EG == 1 || EG == 2, "one or two epilogue groups");
  static_assert(NT % TBUF == 0 && PRE < GT, "tile counts");
  static_assert(HS == 4, "the epilogue is written for 4 samples per warp");
  static_assert(NS % 16 == 0 && TN % 16 == 0 && NS % ST == 0, "cta_group::2 MMAs take N in steps of 16");
  static_assert(WCOL + 32 <= 512, "TMEM budget");
  static_assert(X_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "swizzle atoms need 1024 B alignment");
  static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
  static_assert(XSTAGES <= 4 && ASTAGES <= 8, "barrier slots");
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
    k_body_res(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_d,
               const __grid_constant__ CUtensorMap map_d32, const __grid_constant__ CUtensorMap map_a, const uint4 *__restrict__ W16, int B, int V, int Vp, int n_vp,
               int n_m, float inv_scale, float *__restrict__ verts) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + C::SM_BAR;
  const uint32_t full_d = bar0 + 0;                            // 12: one per (plane, k-block) of the resident tile
  const uint32_t d_empty = bar0 + 96;                          // the blend MMAs of a vertex tile are done
  const uint32_t p_full = bar0 + 104, p_empty = bar0 + 112, w_ready = bar0 + 120;
  const uint32_t full_x = bar0 + 128, empty_x = bar0 + 160;    // C::XSTAGES (<= 4) each
  const uint32_t t_full = bar0 + 192, t_empty = bar0 + 208;    // C::TBUF (2) each
  const uint32_t full_a = bar0 + 320, empty_a = bar0 + 384;    // C::ASTAGES (<= 8) each
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + C::SM_BAR + 288);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;
  // Super-tile t of this pair -> vertex-tile pair t / n_m, sample block t % n_m; this CTA's vertex
  // tile is 2 * (t / n_m) + crank.
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int total = n_vp * n_m;
  const int t0 = (int)(((long long)pair * total) / n_pairs);
  const int t1 = (int)(((long long)(pair + 1) * total) / n_pairs);
  // Walk order inside [t0, t1): ROTATED so that all pairs sweep the sample blocks in step.  Walked from t0,
  // the ~n_pairs / n_vp pairs that share a vertex tile are at unrelated sample blocks and, over all vertex
  // tiles, every sample block of the batch is being written at any time: 3 KB segments scattered over the
  // whole 340 MB -- a store-only kernel with that schedule reaches 3.8 TB/s, with this one 4.5 TB/s
  // (tools/micro/store_pattern.cu).  n_sweep "sweep lines" n_m / n_sweep apart move through the batch;
  // a pair starts at the first of its tiles that lies on one (its range is longer than the spacing) and
  // wraps around, so at any time the pairs are within a block or two of the lines.  Costs at most one
  // extra load of the resident Dt16 tile (a range that straddles two vertex tiles is entered in the middle).
  const int n_t = t1 - t0;
  int rot = 0;
  {
    const int n_sweep = max(1, (n_pairs + n_vp / 2) / n_vp);
    for (int i = 0; i < n_t; ++i) {
      const int m = (t0 + i) % n_m;
      bool on_line = false;
      for (int j = 0; j < n_sweep; ++j) on_line |= m == (j * n_m + n_sweep - 1) / n_sweep;
      if (on_line) {
        rot = i;
        break;
      }
    }
  }
#define TILE_AT(i) (t0 + (((i) + rot) >= n_t ? (i) + rot - n_t : (i) + rot))

  if (threadIdx.x == 0) {
    // (the full_*, p_empty, t_empty and w_ready barriers are only used in the leader)
    for (int i = 0; i < 12; ++i) mbar_init(full_d + 8 * i, 1);
    mbar_init(d_empty, 1);
    mbar_init(p_full, 1);
    mbar_init(p_empty, 8 * C::EW * C::EG);  // one arrival per epilogue warp of the pair
    mbar_init(w_ready, 8 * C::EW);
    for (int i = 0; i < C::XSTAGES; ++i) {
      mbar_init(full_x + 8 * i, 1);
      mbar_init(empty_x + 8 * i, 1);
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
    // =========================== A16 ring producer (both CTAs) ===========================
    if (lane == 0) {
      const uint32_t l_full_a = cluster_map_shared(full_a, 0);
      int stage = 0, phase = 0;
      for (int ti = 0; ti < n_t; ++ti) {
        const int t = TILE_AT(ti);
        const int m = t % n_m;
#pragma unroll 1
        for (int st = 0; st < C::NT; ++st) {
          mbar_wait(empty_a + 8 * stage, phase ^ 1);
          if (FB_ABL2 & 16) {
            if (leader) mbar_arrive(full_a + 8 * stage);
          } else {
            if (leader) mbar_expect_tx(full_a + 8 * stage, 2 * C::A_BYTES);
            tma_load_2d_pair(sbase + C::SM_A + stage * C::A_BYTES, &map_a, 0, (m * C::NS + st * C::ST) * 12 + crank * (C::TN / 2),
                             l_full_a + 8 * stage);
          }
          if (++stage == C::ASTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      // the leader's last "stage free" arrivals must have landed before this CTA may exit
      for (int i = 0; i < C::ASTAGES; ++i) {
        mbar_wait(empty_a + 8 * stage, phase ^ 1);
        if (++stage == C::ASTAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // =========================== blend operand producer: resident Dt16 tile + x16 ring (both CTAs) ===========================
    if (lane == 0) {
      const uint32_t l_full_d = cluster_map_shared(full_d, 0), l_full_x = cluster_map_shared(full_x, 0);
      int cur_vp = -1, loads = 0, stage = 0, phase = 0;
      for (int ti = 0; ti < n_t; ++ti) {
        const int t = TILE_AT(ti);
        const int vp = t / n_m, m = t % n_m;
        if (vp != cur_vp) {
          cur_vp = vp;
          const int vt = 2 * vp + crank;
          if (loads > 0) mbar_wait(d_empty, (loads - 1) & 1);   // every blend MMA on the previous tile has completed
          for (int kb = 0; kb < 4; ++kb)                         // k-block major, the order the blend consumes them in
            for (int cc = 0; cc < 3; ++cc) {
              const int i = cc * 4 + kb;
              if (kb < 3) {
                if (leader) mbar_expect_tx(full_d + 8 * i, 2 * FB_D_BYTES);
                tma_load_2d_pair(sbase + C::SM_D + C::d_off(cc, kb), &map_d, kb * 64, cc * Vp + vt * FB_VT, l_full_d + 8 * i);
              } else {
                if (leader) mbar_expect_tx(full_d + 8 * i, 2 * 3 * C::D_SLAB);
                for (int k = 0; k < 3; ++k)
                  tma_load_2d_pair(sbase + C::SM_D + C::d_off(cc, 3) + k * C::D_SLAB, &map_d32, 192 + 16 * k, cc * Vp + vt * FB_VT,
                                   l_full_d + 8 * i);
              }
            }
          ++loads;
        }
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(empty_x + 8 * stage, phase ^ 1);
          if (FB_ABL2 & 16) {
            if (leader) mbar_arrive(full_x + 8 * stage);
          } else {
            if (leader) mbar_expect_tx(full_x + 8 * stage, 2 * C::X_BYTES);
            tma_load_2d_pair(sbase + C::SM_X + stage * C::X_BYTES, &map_x, kb * 64, m * C::NS + crank * (C::NS / 2),
                             l_full_x + 8 * stage);
          }
          if (++stage == C::XSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (loads > 0) mbar_wait(d_empty, (loads - 1) & 1);      // (also: the last commit's arrival has landed)
      for (int i = 0; i < C::XSTAGES; ++i) {
        mbar_wait(empty_x + 8 * stage, phase ^ 1);
        if (++stage == C::XSTAGES) {
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
      const uint64_t desc_d32 = umma_desc_sw32(sbase + C::SM_D + C::D_MAIN);
      int cur_vp = -1, d_loads = 0, stage = 0, phase = 0, n_tiles = 0;
      [[maybe_unused]] long long w_pe = 0, w_fr = 0, w_fd = 0, w_tot = TCLK(), tq;
      for (int ti = 0; ti < n_t; ++ti, ++n_tiles) {
        const int t = TILE_AT(ti);
        const int vp = t / n_m;
        const bool new_tile = vp != cur_vp;
        if (new_tile) {
          cur_vp = vp;
          ++d_loads;
        }
        tq = TCLK();
        mbar_wait(p_empty, (n_tiles & 1) ^ 1);       // both epilogues have read the previous P
        TADD(w_pe, tq);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          tq = TCLK();
          mbar_wait(full_x + 8 * stage, phase);
          TADD(w_fr, tq);
          const uint64_t b_desc = umma_desc_add(desc_x0, stage * C::X_BYTES);
#pragma unroll 1
          for (int cc = 0; cc < 3; ++cc) {
            if (new_tile) {
              tq = TCLK();
              mbar_wait(full_d + 8 * (cc * 4 + kb), (d_loads - 1) & 1);
              TADD(w_fd, tq);
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + cc * C::NS;
            // k-blocks 0..2: K steps 32 B apart inside a 128-byte swizzle atom; k-block 3: one 4 KB slab per K step
            const uint64_t a_desc = kb < 3 ? umma_desc_add(desc_d0, (cc * 3 + kb) * FB_D_BYTES) : umma_desc_add(desc_d32, cc * 3 * C::D_SLAB);
            const uint32_t a_step = kb < 3 ? 2 : (C::D_SLAB >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kb == 3 && k == 3) continue;      // K = 240
                if (FB_ABLATE != 7 && !(FB_ABL2 & 1)) tc_mma_f16_pair(d_tmem, a_desc + a_step * k, b_desc + 2 * k, idesc_p, (kb | k) != 0);
              }
            }
            __syncwarp();
          }
          if (elect_one()) tc_commit_pair(empty_x + 8 * stage);
          __syncwarp();
          if (++stage == C::XSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        const bool last_of_tile = (ti + 1 == n_t) || (TILE_AT(ti + 1) / n_m != vp);
        if (elect_one()) {
          tc_commit_pair(p_full);
          if (last_of_tile) tc_commit_pair(d_empty);
        }
        __syncwarp();
      }
#ifdef FB_TIMING
      if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 76))
        printf("cta %d blend MMA warp: total %lld, wait p_empty %lld, ring %lld, Dt16 tile %lld (tiles %d)\n", blockIdx.x,
               clock64() - w_tot, w_pe, w_fr, w_fd, n_t);
#endif
    }
  } else if (warp == 3) {
    // =========================== skinning MMA issuer (leader; A = W16 in tensor memory) ===========================
    if (leader) {
      constexpr uint32_t idesc_t = umma_idesc_f16(2 * FB_VT, C::TN);
      const uint64_t desc_a0 = umma_desc_sw128(sbase + C::SM_A);
      const uint32_t w_tmem = tmem_base + C::WCOL;     // window w of a W16 row = columns 8 w .. 8 w + 7
      int cur_vp = -1, w_loads = 0, stage = 0, phase = 0, tb = 0, tphase = 0;
      [[maybe_unused]] long long w_te = 0, w_fa = 0, w_tot = TCLK(), tq;
      for (int ti = 0; ti < n_t; ++ti) {
        const int t = TILE_AT(ti);
        const int vp = t / n_m;
        if (vp != cur_vp) {
          mbar_wait(w_ready, w_loads & 1);             // both epilogues have put this vertex tile's W16 rows in TMEM
          ++w_loads;
          cur_vp = vp;
          tc_fence_after();
        }
#pragma unroll 1
        for (int st = 0; st < C::NT; ++st) {
          tq = TCLK();
          mbar_wait(t_empty + 8 * tb, tphase ^ 1);
          TADD(w_te, tq);
          tq = TCLK();
          mbar_wait(full_a + 8 * stage, phase);
          TADD(w_fa, tq);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + C::TCOL + tb * C::TN;
          const uint64_t a_desc = umma_desc_add(desc_a0, stage * C::A_BYTES);
          if (elect_one()) {
            // (W window, A window) pairs of the table in k_skin_tc.cu; an A window is 32 B = 2 descriptor units
            if (FB_ABLATE != 7 && !(FB_ABL2 & 1)) {   // (tuning builds without MMAs: hand-shakes, TMA traffic and stores only)
              tc_mma_f16_ts_pair(d_tmem, w_tmem + 0, a_desc + 0, idesc_t, 0);
              tc_mma_f16_ts_pair(d_tmem, w_tmem + 8, a_desc + 2, idesc_t, 1);
              tc_mma_f16_ts_pair(d_tmem, w_tmem + 0, a_desc + 4, idesc_t, 1);
              tc_mma_f16_ts_pair(d_tmem, w_tmem + 16, a_desc + 0, idesc_t, 1);
              tc_mma_f16_ts_pair(d_tmem, w_tmem + 24, a_desc + 2, idesc_t, 1);
            }
            tc_commit_pair(empty_a + 8 * stage);
            tc_commit_pair(t_full + 8 * tb);
          }
          __syncwarp();
          if (++stage == C::ASTAGES) {
            stage = 0;
            phase ^= 1;
          }
          if (++tb == C::TBUF) {
            tb = 0;
            tphase ^= 1;
          }
        }
      }
#ifdef FB_TIMING
      if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 76))
        printf("cta %d skin MMA warp: total %lld, wait t_empty %lld, ring %lld\n", blockIdx.x, clock64() - w_tot, w_te, w_fa);
#endif
    }
  } else if (warp >= 4) {
    // =========================== epilogue (warps 4..11, both CTAs) ===========================
    // thread = vertex = TMEM lane; two warps per lane quarter, each 4 of a skinning tile's 8 samples; the
    // v_posed of the last PRE tiles is fetched early so that P can be handed back before the super-tile ends.
    const int q = warp & 3;
    const int part = ((warp - 4) >> 2) & 1;
    const int grp = C::EG == 1 ? 0 : (warp - 4) >> 3;
    constexpr int HS = C::HS;
    constexpr int PRE = C::PRE;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    const uint32_t l_p_empty = cluster_map_shared(p_empty, 0), l_t_empty = cluster_map_shared(t_empty, 0);
    const uint32_t l_w_ready = cluster_map_shared(w_ready, 0);
    int tb = grp, tphase = 0, n_tiles = 0, cur_vp = -1;
    [[maybe_unused]] long long w_pf = 0, w_tf = 0, w_ld = 0, w_st = 0, w_tot = TCLK(), tq;
    for (int ti = 0; ti < n_t; ++ti, ++n_tiles) {
      const int t = TILE_AT(ti);
      const int vp = t / n_m, m = t % n_m, vt = 2 * vp + crank;
      const int v0 = vt * FB_VT + 32 * q;
      const bool v_ok = v0 + lane < V;
      // (FB_ABLATE == 4, tuning build: every super-tile writes the rows of sample block 0 -- the same store
      // instructions and bytes, but the lines stay in L2: no HBM write traffic)
      float *const vbase = verts + ((size_t)(((FB_ABLATE == 4 || (FB_ABL2 & 8)) ? 0 : m) * C::NS) * V + v0 + lane) * 3;
      const int b_left = (FB_ABLATE == 4 || (FB_ABL2 & 8)) ? C::NS : B - m * C::NS;   // samples of this super-tile inside the batch
      if (vp != cur_vp && grp == C::EG - 1) {
        // Every skinning MMA of the previous vertex tile has completed (this warp's group has waited for the
        // T of the super-tile's LAST tile): copy this thread's half of its W16 row (rows >= V are zero) to
        // tensor memory.
        cur_vp = vp;
        uint32_t wr[16];
        const uint4 *src = W16 + (size_t)(v0 + lane) * 8 + part * 4;   // 128 B per row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 x = __ldg(src + i);
          wr[4 * i + 0] = x.x;
          wr[4 * i + 1] = x.y;
          wr[4 * i + 2] = x.z;
          wr[4 * i + 3] = x.w;
        }
        tc_st_32x16_r(lane_base + C::WCOL + part * 16, wr);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(l_w_ready);
      }

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
#if FB_ABLATE == 6 || FB_ABLATE == 7 || (FB_ABL2 & 2)
        // tuning build: no TMEM loads (hand-shakes and stores only)
#pragma unroll
        for (int i = 0; i < 12 * HS; ++i) r[i] = tcol0 + i;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
          for (int si = 0; si < HS; ++si) pc[cc][si] = s_loc + cc;
        if (true) {
#else
        tc_ld_32x32(tcol0, r);
        tc_ld_32x16(tcol0 + 32, r + 32);
        if (p_in == nullptr) {
          load_p(s_loc, pc);
#endif
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
        if (!(FB_ABL2 & 4) || o[0][0] == 1.2345e-30f)
          store_rows4(vbase + s_loc * (V * 3), V * 3, b_left - s_loc, v_ok, o[0][0], o[0][1], o[0][2], o[1][0], o[1][1], o[1][2],
                      o[2][0], o[2][1], o[2][2], o[3][0], o[3][1], o[3][2]);
        TADD(w_st, tq);
        if (C::EG == 1) {
          if (++tb == C::TBUF) {
            tb = 0;
            tphase ^= 1;
          }
        } else {
          tphase ^= 1;     // this group's T stage comes round once per tile of the group
        }
      };

      tq = TCLK();
      mbar_wait(p_full, n_tiles & 1);
      TADD(w_pf, tq);
      tc_fence_after();
      // this group's tiles: grp, grp + EG, ...
#pragma unroll 1
      for (int i = 0; i < C::GT - PRE; ++i) do_tile(grp + i * C::EG, nullptr, PRE == 0 && i == C::GT - 1);
      if (PRE > 0) {
        uint32_t pre[PRE > 0 ? PRE : 1][3][HS];
#pragma unroll
        for (int i = 0; i < PRE; ++i) load_p((grp + (C::GT - PRE + i) * C::EG) * C::ST + part * HS, pre[i]);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(l_p_empty);
#pragma unroll
        for (int i = 0; i < PRE; ++i) do_tile(grp + (C::GT - PRE + i) * C::EG, pre[i], false);
      }
    }
#ifdef FB_TIMING
    if (lane == 0 && (warp == 4 || warp == 9 || warp == 12) && (blockIdx.x == 0 || blockIdx.x == 1 || blockIdx.x == 76))
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
using ResA = ResCfg<96, 8, 3, 4, 2, 1>;   // eight epilogue warps; rings: x16 3 stages, A16 4 (the default)
using ResB = ResCfg<96, 8, 2, 3, 1, 2>;   // sixteen epilogue warps in two groups
using ResC = ResCfg<96, 8, 2, 3, 2, 1>;   // round 2's first ring depths (2 / 3), for comparison

int body_res_init(smplb_ctx *c) {
  CUDA_TRY(cudaFuncSetAttribute(k_body_res<ResA>, cudaFuncAttributeMaxDynamicSharedMemorySize, ResA::SM_TOTAL));
  CUDA_TRY(cudaFuncSetAttribute(k_body_res<ResB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ResB::SM_TOTAL));
  CUDA_TRY(cudaFuncSetAttribute(k_body_res<ResC>, cudaFuncAttributeMaxDynamicSharedMemorySize, ResC::SM_TOTAL));
  // the last k-block of the resident tile as 16-column SWIZZLE_32B boxes (same rows as map_d)
  TRY(tc_make_map(c->map_d32, 0, c->d_Dt16, 256, (uint64_t)c->pitch, 512, 16, FB_VT, /*swizzle=*/32));
  return 0;
}

template <class C>
static int launch_res_cfg(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts) {
  alignas(64) CUtensorMap map_x, map_a;
  TRY(tc_make_map(&map_x, 0, x16, 256, (uint64_t)B, 512, 64, C::NS / 2));
  TRY(tc_make_map(&map_a, 0, A16, 64, (uint64_t)B * 12, 128, 64, C::TN / 2));
  const int n_vp = c->Vp / (2 * FB_VT), n_m = cdiv(B, C::NS);
  const int total = n_vp * n_m;
  const int want_pairs = c->body_pairs != 0 ? c->body_pairs : c->pairs_auto;   // (-1 / 0: every SM pair)
  const int max_pairs = want_pairs > 0 && want_pairs < c->num_sms / 2 ? want_pairs : c->num_sms / 2;
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
  const CUtensorMap map_d = *(const CUtensorMap *)c->map_d, map_d32 = *(const CUtensorMap *)c->map_d32;
  const uint4 *W16 = (const uint4 *)c->d_W16;
  int Vv = c->V, Vp = c->Vp;
  float inv = c->tc_inv_scale;
  {
    ProfScope ps(c, "body_fwd_tc");
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_body_res<C>, map_x, map_d, map_d32, map_a, W16, B, Vv, Vp, n_vp, n_m, inv, verts);
    if (e != cudaSuccess) {
      smplb_set_error("launch body_fwd_tc (resident Dt16 tile) failed: %s", cudaGetErrorString(e));
      return SMPLB_ECUDA;
    }
  }
  c->launches++;
  return 0;
}

// verts [B][V][3]; needs an even number of 128-vertex tiles (the caller checks).  variant 9: sixteen epilogue warps in
// two groups; variant 10: the 2 / 3 ring depths.
int launch_body_fwd_res(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts, int variant) {
  if (variant == 9) return launch_res_cfg<ResB>(c, B, x16, A16, verts);
  if (variant == 10) return launch_res_cfg<ResC>(c, B, x16, A16, verts);
  return launch_res_cfg<ResA>(c, B, x16, A16, verts);
}
