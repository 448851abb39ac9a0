// Generic fp16 x fp16 (or bf16 x bf16) -> fp32 GEMM on tcgen05/TMEM, TMA-fed:
//
//   C[z][m, n] = scale * sum_{k in split z} A[m, k] * B[n, k]        (both operands K-major)
//
// Used for the two contractions of the folded keypoint path (k_fold.cu) and, with bf16 operands, for the
// blend-transpose contraction of the dense backward (d pose_feature = dp . posedirs^T, K = 3 * 20736).  Split-precision is
// expressed by the caller along K (operands concatenated as hi|hi|lo against hi|lo|hi), so this
// kernel is a plain GEMM.  Persistent, warp-specialised (TMA producer / MMA issuer / 4 epilogue
// warps), 128 x 128 output tiles, a 4-deep ring of 64-wide k-blocks, two TMEM accumulator
// stages, epilogue through swizzled shared memory and TMA stores (rows beyond M are clipped).
#include <cuda.h>
#include <cuda_fp16.h>

#include "smplb_internal.h"
#include "tc_ptx.cuh"

#define G_BM 128
#define G_BN 128
#define G_KB 64
#define G_STAGES 4
#define G_THREADS 192
#define G_KB_BYTES (128 * 128)                 // one operand k-block: 128 rows x 128 B
#define G_STAGE_BYTES (2 * G_KB_BYTES)         // A + B
#define G_SM_C (G_STAGES * G_STAGE_BYTES)      // 128 KB
#define G_SM_BAR (G_SM_C + 4 * 8192)
#define G_SM_TOTAL (G_SM_BAR + 128)

__global__ void __launch_bounds__(G_THREADS, 1)
    k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const __grid_constant__ CUtensorMap map_c, int n_mblk, int n_nblk, int n_kblk, int ksplit, int c_rows_per_split,
              float scale, uint32_t idesc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + G_SM_BAR;
  const uint32_t full = bar0 + 0, empty = bar0 + 32, tmem_full = bar0 + 64, tmem_empty = bar0 + 80;
  volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + G_SM_BAR + 96);
  // (the shuffle tells the compiler the warp index is warp-uniform: role branches and the addresses
  // derived from it stay in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int total = n_mblk * n_nblk * ksplit;
  const int kb_per = (n_kblk + ksplit - 1) / ksplit;

  if (threadIdx.x == 0) {
    for (int i = 0; i < G_STAGES; ++i) {
      mbar_init(full + 8 * i, 1);
      mbar_init(empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + 8 * i, 1);
      mbar_init(tmem_empty + 8 * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + G_SM_BAR + 96),
                 "n"(2 * G_BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile t -> (z, m, n), n fastest so consecutive tiles of a CTA reuse the A rows from L2
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int n = t % n_nblk, m = (t / n_nblk) % n_mblk, z = t / (n_nblk * n_mblk);
        int kb0 = z * kb_per, kb1 = min(n_kblk, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty + 8 * stage, phase ^ 1);
          mbar_expect_tx(full + 8 * stage, G_STAGE_BYTES);
          tma_load_2d(sbase + stage * G_STAGE_BYTES, &map_a, kb * G_KB, m * G_BM, full + 8 * stage);
          tma_load_2d(sbase + stage * G_STAGE_BYTES + G_KB_BYTES, &map_b, kb * G_KB, n * G_BN, full + 8 * stage);
          if (++stage == G_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // all 32 lanes run the loop, the elected lane issues (tc_ptx.cuh: elect_one)
    {
      const uint64_t desc0 = umma_desc_sw128(sbase);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int z = t / (n_nblk * n_mblk);
        int kb0 = z * kb_per, kb1 = min(n_kblk, kb0 + kb_per);
        mbar_wait(tmem_empty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        uint32_t d_tmem = tmem_base + acc * G_BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_add(desc0, stage * G_STAGE_BYTES), b_desc = umma_desc_add(a_desc, G_KB_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < G_KB / 16; ++k) tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb != kb0) || (k != 0));
            tc_commit(empty + 8 * stage);
          }
          __syncwarp();
          if (++stage == G_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) tc_commit(tmem_full + 8 * acc);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const int q = warp & 3;
    const uint32_t stage_base = sbase + G_SM_C + (warp - 2) * 8192;
    int acc = 0, acc_phase = 0, buf = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int n = t % n_nblk, m = (t / n_nblk) % n_mblk, z = t / (n_nblk * n_mblk);
      int kb0 = z * kb_per, kb1 = min(n_kblk, kb0 + kb_per);
      mbar_wait(tmem_full + 8 * acc, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int cch = 0; cch < G_BN / 32; ++cch) {
        uint32_t r[32];
        tc_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * G_BN + cch * 32, r);
        tc_wait_ld();
        if (cch == G_BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + 8 * acc);
        }
        if (lane == 0) tma_wait_read<1>();
        __syncwarp();
        uint32_t row_addr = stage_base + buf * 4096 + lane * 128;
        const float sc = (kb1 > kb0) ? scale : 0.0f;   // an empty K range contributes zeros
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float x = __uint_as_float(r[4 * j + 0]) * sc, y = __uint_as_float(r[4 * j + 1]) * sc;
          float zz = __uint_as_float(r[4 * j + 2]) * sc, w = __uint_as_float(r[4 * j + 3]) * sc;
          if (kb1 <= kb0) x = y = zz = w = 0.0f;
          uint32_t addr = row_addr + ((j ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(zz), "f"(w) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&map_c, stage_base + buf * 4096, n * G_BN + cch * 32, z * c_rows_per_split + m * G_BM + 32 * q);
          tma_commit();
        }
        buf ^= 1;
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * G_BN) : "memory");
  }
}

// The one place tensor maps are encoded (cuTensorMapEncodeTiled through the runtime's driver entry
// point, so libsmplb.so does not link libcuda): 2-D row-major tensors of fp16 (is_f32 = 0) or fp32
// elements, box = box_inner x box_outer elements, 128-byte swizzle unless swizzle = 0.
typedef CUresult (*encode_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_fn_t g_enc = nullptr;

int tc_make_map(void *map, int is_f32, const void *ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                uint32_t box_inner, uint32_t box_outer, int swizzle) {
  if (!g_enc) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    RET_IF(!fn || qres != cudaDriverEntryPointSuccess, SMPLB_ECUDA, "cuTensorMapEncodeTiled is unavailable");
    g_enc = (encode_fn_t)fn;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_enc((CUtensorMap *)map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RET_IF(r != CUDA_SUCCESS, SMPLB_ECUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

// C [ksplit][c_rows_per_split][ldc] fp32 = scale * A16 [M][K] * B16 [N][K]^T.  K is a multiple of
// 64; N is padded by the tensor map's zero fill; c_rows_per_split is M rounded up to 128.
// bf16 != 0: both operands are bfloat16 (same bytes per element, same rate; the TMA descriptor only moves bytes).
int launch_gemm_tc(smplb_ctx *c, const char *name, int M, int N, int K, const void *A16, const void *map_b, float *C,
                   int ldc, int ksplit, float scale, int bf16) {
  if (!(c->attr_done & 2u)) {
    CUDA_TRY(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SM_TOTAL));
    c->attr_done |= 2u;
  }
  int n_mblk = cdiv(M, G_BM), n_nblk = cdiv(N, G_BN), n_kblk = K / G_KB;
  int rows_per = n_mblk * G_BM;
  alignas(64) CUtensorMap map_a, map_c;
  TRY(tc_make_map(&map_a, 0, (void *)A16, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, G_KB, G_BM));
  // one split: clip the stores at the real M; several: the caller's buffer holds rows_per rows per split
  uint64_t c_rows = ksplit == 1 ? (uint64_t)M : (uint64_t)ksplit * rows_per;
  TRY(tc_make_map(&map_c, 1, (void *)C, (uint64_t)ldc, c_rows, (uint64_t)ldc * 4, 32, 32));
  int total = n_mblk * n_nblk * ksplit;
  int grid = total < c->num_sms ? total : c->num_sms;
  // instruction descriptor (tc_ptx.cuh); a_format / b_format = 1 (bits 7, 10) selects bf16 operands
  const uint32_t idesc = umma_idesc_f16(G_BM, G_BN) | (bf16 ? ((1u << 7) | (1u << 10)) : 0u);
  LAUNCH(c, name, grid, G_THREADS, G_SM_TOTAL, k_gemm_tc, map_a, *(const CUtensorMap *)map_b, map_c, n_mblk, n_nblk, n_kblk,
         ksplit, rows_per, scale, idesc);
  return 0;
}
