// Blend-shape contraction on FP32 CUDA cores (validation / bring-up path; the production
// forward is the tcgen05 kernel in k_blend_tc.cu once enabled).
//
//   forward  (batch_smpl.py:110-112 and :126-132 in one GEMM):
//       v_posed[b, 3v+c] = sum_k x[b,k] * Dext[k, 3v+c],   x = [pose_feature | beta | 1]
//       Dext rows = posedirs (207) | shapedirs (NB) | v_template (1) | zero padding
//   backward (TF autodiff of the same two matmuls):
//       dx[b,k] = sum_n dp[b,n] * Dext[k,n]   (split-K partial sums, reduced in k_pose_bwd)
#include <cuda_bf16.h>

#include "smplb_internal.h"

#define BM 128
#define BN 128
#define BK 8

// C[M, ldc] (+z * M * ldc) = A[M, lda (k contiguous)] * op(Bm)
//   TRANS_B == false: Bm is [K, ldb] with n contiguous          (forward, NN)
//   TRANS_B == true : Bm is [N, ldb] with k contiguous          (backward, NT)
// blockIdx.z selects a K range [z*kchunk, min(K, (z+1)*kchunk)).  K ranges and ld's are
// multiples of 8 / 4 so every global load is an aligned float4.  Rows >= M read as zero.
template <bool TRANS_B>
__global__ void __launch_bounds__(256) k_sgemm(int M, int N, int K, int kchunk, const float *__restrict__ A, int lda,
                                               const float *__restrict__ Bm, int ldb, float *__restrict__ C, int ldc) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  int tid = threadIdx.x;
  int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int kbeg = blockIdx.z * kchunk;
  int kend = min(K, kbeg + kchunk);
  int ty = tid / 16, tx = tid % 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  // A tile loader: 128 rows x 8 k = 256 float4 -> one per thread
  int a_row = tid / 2, a_k4 = (tid % 2) * 4;
  // B tile loader
  int b_k = tid / 32, b_n4 = (tid % 32) * 4;   // NN: 8 k-rows x 32 float4
  int bt_row = tid / 2, bt_k4 = (tid % 2) * 4; // NT: 128 n-rows x 2 float4

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + a_row < M) av = *reinterpret_cast<const float4 *>(A + (size_t)(m0 + a_row) * lda + k0 + a_k4);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!TRANS_B) {
      if (n0 + b_n4 < N) bv = *reinterpret_cast<const float4 *>(Bm + (size_t)(k0 + b_k) * ldb + n0 + b_n4);
    } else {
      if (n0 + bt_row < N) bv = *reinterpret_cast<const float4 *>(Bm + (size_t)(n0 + bt_row) * ldb + k0 + bt_k4);
    }
    __syncthreads();  // previous tile fully consumed
    As[a_k4 + 0][a_row] = av.x;
    As[a_k4 + 1][a_row] = av.y;
    As[a_k4 + 2][a_row] = av.z;
    As[a_k4 + 3][a_row] = av.w;
    if (!TRANS_B) {
      *reinterpret_cast<float4 *>(&Bs[b_k][b_n4]) = bv;
    } else {
      Bs[bt_k4 + 0][bt_row] = bv.x;
      Bs[bt_k4 + 1][bt_row] = bv.y;
      Bs[bt_k4 + 2][bt_row] = bv.z;
      Bs[bt_k4 + 3][bt_row] = bv.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float *Cz = C + (size_t)blockIdx.z * M * ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int n = n0 + h * 64 + tx * 4;
      if (n + 3 < N) {
        *reinterpret_cast<float4 *>(Cz + (size_t)m * ldc + n) =
            make_float4(acc[i][4 * h + 0], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
      } else {
        for (int q = 0; q < 4; ++q)
          if (n + q < N) Cz[(size_t)m * ldc + n + q] = acc[i][4 * h + q];
      }
    }
  }
}

int launch_blend_fwd(smplb_ctx *c, int B, const float *x, float *v_posed) {
  // N = pitch (Dext is zero beyond 3V, so the padding columns of v_posed become 0)
  dim3 grid(c->pitch / BN, cdiv(B, BM), 1);
  LAUNCH(c, "blend_fwd_sgemm", grid, 256, 0, k_sgemm<false>, B, c->pitch, KX, KX, x, KX, c->d_Dext, c->pitch, v_posed,
         c->pitch);
  return 0;
}

int launch_blend_bwd(smplb_ctx *c, int B, const float *dp, float *dx_part, bool compact, int ksplit) {
  // compact: dp holds only the columns of the active vertices (rows of joint_regressor with a
  // non-zero), Dext_act the matching columns of Dext; every skipped column of dp is exactly 0.
  int pitch = compact ? c->pitch_act : c->pitch;
  const float *D = compact ? c->d_Dext_act : c->d_Dext;
  int kchunk = cdiv(pitch / BK, ksplit) * BK;
  dim3 grid(cdiv(KX, BN), cdiv(B, BM), ksplit);
  LAUNCH(c, compact ? "blend_bwd_sgemm_active" : "blend_bwd_sgemm", grid, 256, 0, k_sgemm<true>, B, KX, pitch, kchunk, dp,
         pitch, D, pitch, dx_part, KX);
  return 0;
}

// ---- dense blend backward on the tensor cores ------------------------------------------------------
// dx[b, k] = sum_n dp[b, n] Dext[k, n] is a [B x 62208] x [62208 x 224] contraction once the operands are split
// (dp = hi + lo and Dext = hi + lo in bf16; hi.hi + lo.hi + hi.lo keeps ~16 significand bits, the dropped lo.lo
// term is 2^-16 relative): k_gemm_tc with bf16 operands, split-K over the SMs, partials summed by k_pose_bwd.
__global__ void k_build_dbf(int pitch, const float *__restrict__ Dext, __nv_bfloat16 *__restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (n >= pitch) return;
  const float v = Dext[(size_t)k * pitch + n];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16 *row = out + (size_t)k * 3 * pitch;
  row[n] = hi;
  row[pitch + n] = hi;
  row[2 * (size_t)pitch + n] = lo;
}

int blend_bwd_tc_init(smplb_ctx *c) {
  if (c->blend_bwd_tc_ok) return 0;
  RET_IF(c->pitch % 64 != 0, SMPLB_ESTATE, "pitch is not a multiple of the GEMM's k-block");
  CUDA_TRY(cudaMalloc(&c->d_Dbf, (size_t)KX * 3 * c->pitch * 2));
  k_build_dbf<<<dim3(cdiv(c->pitch, 256), KX), 256, 0, c->cur>>>(c->pitch, c->d_Dext, (__nv_bfloat16 *)c->d_Dbf);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  TRY(tc_make_map(c->map_dbf, 0, c->d_Dbf, (uint64_t)3 * c->pitch, (uint64_t)KX, (uint64_t)3 * c->pitch * 2, 64, 128));
  c->blend_bwd_tc_ok = true;
  return 0;
}

int launch_blend_bwd_tc(smplb_ctx *c, int B, const void *dp16, float *dx_part, int *ksplit, int *dx_rows) {
  TRY(blend_bwd_tc_init(c));
  const int n_mblk = cdiv(B, 128), n_nblk = cdiv(KX, 128);
  int ks = c->num_sms / (n_mblk * n_nblk);
  ks = ks < 1 ? 1 : (ks > 16 ? 16 : ks);
  *ksplit = ks;
  *dx_rows = n_mblk * 128;
  return launch_gemm_tc(c, "blend_bwd_tc", B, KX, 3 * c->pitch, dp16, c->map_dbf, dx_part, KX, ks, 1.0f, /*bf16=*/1);
}
