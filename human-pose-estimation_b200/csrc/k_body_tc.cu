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
// k_skin_tc.cu).  A tcgen05.mma costs >= ~100 clk whatever its N (tools/micro/tmem_rate.cu), so the
// shapes are chosen to keep the instruction count per sample low under the 512-column TMEM
// budget: 3 * NS (P) + TBUF * 12 * ST (T) <= 512.
//
// Persistent, warp-specialised: warp 0 TMA producer of the blend operands, warp 1 MMA issuer and
// TMEM allocator, warps 2-9 epilogue (two per TMEM lane quarter, each half of a tile's samples),
// warp 10 TMA producer of the skinning operands.
#include <cuda.h>
#include <cuda_fp16.h>

#include "smplb_internal.h"
#include "tc_ptx.cuh"

#ifndef FB_NS
#define FB_NS 96                      // samples per super-tile (blend MMA N)
#endif
#ifndef FB_ST
#define FB_ST 8                       // samples per skinning MMA
#endif
#ifndef FB_TBUF
#define FB_TBUF 2                     // T accumulator stages
#endif
#define FB_VT 128                     // vertices per super-tile (MMA M)
#define FB_TN (12 * FB_ST)            // skinning MMA N
#define FB_NT (FB_NS / FB_ST)         // skinning tiles per super-tile
#define FB_DSTAGES 4
#define FB_ASTAGES 3
#define FB_THREADS 352
#define FB_X_KB_BYTES (FB_NS * 128)                 // one k-block of the x16 tile
#define FB_X_BYTES (4 * FB_X_KB_BYTES)
#define FB_D_BYTES (FB_VT * 128)                    // one Dt16 k-block of one plane: 16 KB
#define FB_W_BYTES (FB_VT * 128)                    // W16 tile: 16 KB
#define FB_A_BYTES (FB_TN * 128)                    // A16 rows of ST samples
#define FB_SM_X 0
#define FB_SM_D (FB_SM_X + FB_X_BYTES)
#define FB_SM_W (FB_SM_D + FB_DSTAGES * FB_D_BYTES)
#define FB_SM_A (FB_SM_W + 2 * FB_W_BYTES)
#define FB_SM_BAR (FB_SM_A + FB_ASTAGES * FB_A_BYTES)
#define FB_SM_TOTAL (FB_SM_BAR + 256)
#define FB_TCOL (3 * FB_NS)                         // first TMEM column of the T stages

static_assert(FB_NS % 16 == 0 && FB_NS % FB_ST == 0 && FB_ST % 8 == 0, "tile shape");
static_assert(3 * FB_NS + FB_TBUF * FB_TN <= 512, "TMEM budget");
static_assert(FB_X_KB_BYTES % 1024 == 0 && FB_A_BYTES % 1024 == 0, "swizzle atoms need 1024 B alignment");
static_assert(FB_SM_TOTAL <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(FB_THREADS, 1)
    k_body_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_d,
              const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_a, int B, int V, int Vp,
              int n_vt, int n_m, float inv_scale, float *__restrict__ verts) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + FB_SM_BAR;
  const uint32_t full_x = bar0 + 0, empty_x = bar0 + 8, p_full = bar0 + 16, p_empty = bar0 + 24;
  const uint32_t full_d = bar0 + 32, empty_d = bar0 + 64;     // FB_DSTAGES (<= 4) each
  const uint32_t full_w = bar0 + 96, empty_w = bar0 + 112;    // 2 each
  const uint32_t full_a = bar0 + 128, empty_a = bar0 + 160;   // FB_ASTAGES (<= 4) each
  const uint32_t t_full = bar0 + 192, t_empty = bar0 + 208;   // FB_TBUF (<= 2) each
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + FB_SM_BAR + 224);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = n_vt * n_m;
  const int t0 = (int)(((long long)blockIdx.x * total) / gridDim.x);
  const int t1 = (int)(((long long)(blockIdx.x + 1) * total) / gridDim.x);

  if (threadIdx.x == 0) {
    mbar_init(full_x, 1);
    mbar_init(empty_x, 1);
    mbar_init(p_full, 1);
    mbar_init(p_empty, 8);                 // one arrival per epilogue warp
    for (int i = 0; i < FB_DSTAGES; ++i) {
      mbar_init(full_d + 8 * i, 1);
      mbar_init(empty_d + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(full_w + 8 * i, 1);
      mbar_init(empty_w + 8 * i, 1);
    }
    for (int i = 0; i < FB_ASTAGES; ++i) {
      mbar_init(full_a + 8 * i, 1);
      mbar_init(empty_a + 8 * i, 1);
    }
    for (int i = 0; i < FB_TBUF; ++i) {
      mbar_init(t_full + 8 * i, 1);
      mbar_init(t_empty + 8 * i, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + FB_SM_BAR + 224), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =========================== blend operand producer ===========================
    // Super-tiles are ordered sample-block major, so the x16 tile is loaded once per sample
    // block; the Dt16 k-blocks of the vertex tile (3 planes x 4 k-blocks, L2-resident) stream
    // through the ring.
    if (lane == 0) {
      int cur_m = -1, x_loads = 0, stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int m = t / n_vt, vt = t % n_vt;
        if (m != cur_m) {
          if (x_loads > 0) mbar_wait(empty_x, (x_loads - 1) & 1);
          mbar_expect_tx(full_x, FB_X_BYTES);
          for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + FB_SM_X + kb * FB_X_KB_BYTES, &map_x, kb * 64, m * FB_NS, full_x);
          ++x_loads;
          cur_m = m;
        }
        for (int cc = 0; cc < 3; ++cc)
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(empty_d + 8 * stage, phase ^ 1);
            mbar_expect_tx(full_d + 8 * stage, FB_D_BYTES);
            tma_load_2d(sbase + FB_SM_D + stage * FB_D_BYTES, &map_d, kb * 64, cc * Vp + vt * FB_VT, full_d + 8 * stage);
            if (++stage == FB_DSTAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
      }
    }
  } else if (warp == 10) {
    // =========================== skinning operand producer ===========================
    if (lane == 0) {
      int wbuf = 0, wphase = 0, stage = 0, phase = 0;
      for (int t = t0; t < t1; ++t) {
        const int m = t / n_vt, vt = t % n_vt;
        mbar_wait(empty_w + 8 * wbuf, wphase ^ 1);
        mbar_expect_tx(full_w + 8 * wbuf, FB_W_BYTES);
        tma_load_2d(sbase + FB_SM_W + wbuf * FB_W_BYTES, &map_w, 0, vt * FB_VT, full_w + 8 * wbuf);
        if (++wbuf == 2) {
          wbuf = 0;
          wphase ^= 1;
        }
        for (int st = 0; st < FB_NT; ++st) {
          mbar_wait(empty_a + 8 * stage, phase ^ 1);
          mbar_expect_tx(full_a + 8 * stage, FB_A_BYTES);
          tma_load_2d(sbase + FB_SM_A + stage * FB_A_BYTES, &map_a, 0, (m * FB_NS + st * FB_ST) * 12, full_a + 8 * stage);
          if (++stage == FB_ASTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc_p = umma_idesc_f16(FB_VT, FB_NS);
      constexpr uint32_t idesc_t = umma_idesc_f16(FB_VT, FB_TN);
      int cur_m = -1, x_loads = 0, dstage = 0, dphase = 0, wbuf = 0, wphase = 0, astage = 0, aphase = 0;
      int tb = 0, tphase = 0, n_tiles = 0;
      for (int t = t0; t < t1; ++t, ++n_tiles) {
        const int m = t / n_vt;
        if (m != cur_m) {
          mbar_wait(full_x, x_loads & 1);
          ++x_loads;
          cur_m = m;
        }
        // ---- blend: P = Dt16 tile x x16 tile^T, three coordinate planes
        mbar_wait(p_empty, (n_tiles & 1) ^ 1);       // the epilogue has read the previous P
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 3; ++cc) {
          const uint32_t d_tmem = tmem_base + cc * FB_NS;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(full_d + 8 * dstage, dphase);
            tc_fence_after();
            const uint32_t a_addr = sbase + FB_SM_D + dstage * FB_D_BYTES;
            const uint32_t b_addr = sbase + FB_SM_X + kb * FB_X_KB_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kb == 3 && k == 3) continue;        // K = 240: the last 16 columns are zero padding
              tc_mma_f16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc_p, (kb | k) != 0);
            }
            tc_commit(empty_d + 8 * dstage);
            if (++dstage == FB_DSTAGES) {
              dstage = 0;
              dphase ^= 1;
            }
          }
        }
        tc_commit(p_full);
        const bool last_of_m = (t + 1 == t1) || ((t + 1) / n_vt != m);
        if (last_of_m) tc_commit(empty_x);
        // ---- skinning transforms, ST samples at a time
        mbar_wait(full_w + 8 * wbuf, wphase);
        const uint32_t w_addr = sbase + FB_SM_W + wbuf * FB_W_BYTES;
#pragma unroll 1
        for (int st = 0; st < FB_NT; ++st) {
          mbar_wait(t_empty + 8 * tb, tphase ^ 1);
          mbar_wait(full_a + 8 * astage, aphase);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + FB_TCOL + tb * FB_TN;
          const uint32_t a_addr = sbase + FB_SM_A + astage * FB_A_BYTES;
          // (W window, A window) pairs of the table in k_skin_tc.cu
          tc_mma_f16(d_tmem, umma_desc_sw128(w_addr + 0 * 32), umma_desc_sw128(a_addr + 0 * 32), idesc_t, 0);
          tc_mma_f16(d_tmem, umma_desc_sw128(w_addr + 1 * 32), umma_desc_sw128(a_addr + 1 * 32), idesc_t, 1);
          tc_mma_f16(d_tmem, umma_desc_sw128(w_addr + 0 * 32), umma_desc_sw128(a_addr + 2 * 32), idesc_t, 1);
          tc_mma_f16(d_tmem, umma_desc_sw128(w_addr + 2 * 32), umma_desc_sw128(a_addr + 0 * 32), idesc_t, 1);
          tc_mma_f16(d_tmem, umma_desc_sw128(w_addr + 3 * 32), umma_desc_sw128(a_addr + 1 * 32), idesc_t, 1);
          tc_commit(empty_a + 8 * astage);
          tc_commit(t_full + 8 * tb);
          if (++astage == FB_ASTAGES) {
            astage = 0;
            aphase ^= 1;
          }
          if (++tb == FB_TBUF) {
            tb = 0;
            tphase ^= 1;
          }
        }
        tc_commit(empty_w + 8 * wbuf);
        if (++wbuf == 2) {
          wbuf = 0;
          wphase ^= 1;
        }
      }
    }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    // Two warps per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31); each takes
    // half of a skinning tile's samples.  Thread = vertex: T (12 columns per sample) and the three
    // coordinates of v_posed come out of TMEM; nothing passes through shared memory.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HS = FB_ST / 2;                 // samples per warp per tile
    int tb = 0, tphase = 0, n_tiles = 0;
    for (int t = t0; t < t1; ++t, ++n_tiles) {
      const int m = t / n_vt, vt = t % n_vt;
      const int v0 = vt * FB_VT + 32 * q;
      const bool v_ok = v0 + lane < V;
      const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
      mbar_wait(p_full, n_tiles & 1);
#pragma unroll 1
      for (int st = 0; st < FB_NT; ++st) {
        mbar_wait(t_full + 8 * tb, tphase);
        tc_fence_after();
        const int s_loc = st * FB_ST + half * HS;   // first sample (within the super-tile) of this warp
#pragma unroll
        for (int g = 0; g < HS / 4; ++g) {
          uint32_t r[48], pc[3][4];
          const uint32_t tcol = lane_base + FB_TCOL + tb * FB_TN + (half * HS + 4 * g) * 12;
          tc_ld_32x32(tcol, r);
          tc_ld_32x16(tcol + 32, r + 32);
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) tc_ld_32x4(lane_base + cc * FB_NS + s_loc + 4 * g, pc[cc]);
          tc_wait_ld();
          if (g == HS / 4 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(t_empty + 8 * tb);
              if (st == FB_NT - 1) mbar_arrive(p_empty);   // last read of this super-tile's P
            }
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
            const int b = m * FB_NS + s_loc + 4 * g + si;
            if (b < B && v_ok) {
              float *dst = verts + ((size_t)b * V + v0 + lane) * 3;
              __stcs(dst, o[0]);
              __stcs(dst + 1, o[1]);
              __stcs(dst + 2, o[2]);
            }
          }
        }
        if (++tb == FB_TBUF) {
          tb = 0;
          tphase ^= 1;
        }
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

// ------------------------------------------------------------------------------ host side
typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_fn_t g_encode3 = nullptr;

static int make_map_rows16(CUtensorMap *map, const void *ptr, uint64_t row_halves, uint64_t rows, uint32_t box_rows) {
  if (!g_encode3) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    RET_IF(!fn || qres != cudaDriverEntryPointSuccess, SMPLB_ECUDA, "cuTensorMapEncodeTiled is unavailable");
    g_encode3 = (encode_fn_t)fn;
  }
  cuuint64_t dims[2] = {row_halves, rows};
  cuuint64_t strides[1] = {row_halves * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode3(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RET_IF(r != CUDA_SUCCESS, SMPLB_ECUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

// Needs the operands of both tensor-core kernels (Dt16 + its scale, W16).
int body_tc_init(smplb_ctx *c) {
  c->body_tc_ok = false;
  if (!c->tc_ok || !c->skin_tc_ok) return 0;
  CUDA_TRY(cudaFuncSetAttribute(k_body_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SM_TOTAL));
  c->body_tc_ok = true;
  return 0;
}

// verts [B][V][3] from the operand rows pose_fwd wrote (x16 [B][256], A16 [12 B][64]).
int launch_body_fwd_tc(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts) {
  RET_IF(!c->body_tc_ok, SMPLB_ESTATE, "fused tcgen05 blend+skinning path is not initialised");
  alignas(64) CUtensorMap map_x, map_a;
  TRY(make_map_rows16(&map_x, x16, 256, (uint64_t)B, FB_NS));
  TRY(make_map_rows16(&map_a, A16, 64, (uint64_t)B * 12, FB_TN));
  const int n_vt = c->Vp / FB_VT, n_m = cdiv(B, FB_NS);
  const int total = n_vt * n_m;
  const int grid = total < c->num_sms ? total : c->num_sms;
  LAUNCH(c, "body_fwd_tc", grid, FB_THREADS, FB_SM_TOTAL, k_body_tc, map_x, *(const CUtensorMap *)c->map_d,
         *(const CUtensorMap *)c->map_w, map_a, B, c->V, c->Vp, n_vt, n_m, c->tc_inv_scale, verts);
  return 0;
}
