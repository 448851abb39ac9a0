// SURVEY.md section 8f ("next" rows adjacent to the hot path):
//   * silhouette extraction: tf.cast(tf.where(seg > 0)[:, :3], float32) of src/trainer.py:291,443
//     as an ordered on-device compaction straight into the CSR form the mesh loss consumes;
//   * get_kcs (src/models.py:123-139): KCS = B^T B with B = joints[:, :14]^T C, forward and
//     backward.  The reference builds an N x 13 x 13 x N intermediate to take its diagonal;
//     here it is 169 dot products of length 3 per sample.
#include "smplb_internal.h"

#define FULL 0xffffffffu

// counts[i] = #{seg[i] > 0}
__global__ void __launch_bounds__(256) k_sil_count(int HW, const float *__restrict__ seg, int *__restrict__ counts) {
  __shared__ int red[256];
  int i = blockIdx.x, t = threadIdx.x;
  const float *s = seg + (size_t)i * HW;
  int c = 0;
  for (int k = t; k < HW; k += 256) c += s[k] > 0.0f;
  red[t] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  if (t == 0) counts[i] = red[0];
}

// offsets = exclusive scan of counts (single block, B <= any: sequential chunks of 1024)
__global__ void __launch_bounds__(1024) k_sil_scan(int B, const int *__restrict__ counts, int *__restrict__ offsets) {
  __shared__ int buf[1024];
  __shared__ int carry;
  int t = threadIdx.x;
  if (t == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    int v = (b0 + t < B) ? counts[b0 + t] : 0;
    buf[t] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int x = t >= o ? buf[t - o] : 0;
      __syncthreads();
      buf[t] += x;
      __syncthreads();
    }
    if (b0 + t < B) offsets[b0 + t] = carry + buf[t] - v;
    __syncthreads();
    if (t == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (t == 0) offsets[B] = carry;
}

// points[offsets[i] + rank] = (x = col, y = row) of the rank-th pixel > 0 in row-major order.
__global__ void __launch_bounds__(256) k_sil_fill(int H, int W, const float *__restrict__ seg,
                                                  const int *__restrict__ offsets, int cap, float *__restrict__ points) {
  __shared__ int wsum[8];
  __shared__ int base;
  int i = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  int HW = H * W;
  const float *s = seg + (size_t)i * HW;
  if (t == 0) base = offsets[i];
  __syncthreads();
  for (int k0 = 0; k0 < HW; k0 += 256) {
    int k = k0 + t;
    bool on = k < HW && s[k] > 0.0f;
    unsigned m = __ballot_sync(FULL, on);
    int rank_in_warp = __popc(m & ((1u << lane) - 1));
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int q = 0; q < 8; ++q) {
      if (q < w) before += wsum[q];
      total += wsum[q];
    }
    int pos = base + before + rank_in_warp;
    if (on && pos < cap) {
      points[2 * (size_t)pos + 0] = (float)(k % W);   // x = column  (ops.py:123: silhouette_gt[:, 2])
      points[2 * (size_t)pos + 1] = (float)(k / W);   // y = row     (ops.py:124: silhouette_gt[:, 1])
    }
    __syncthreads();
    if (t == 0) base += total;
    __syncthreads();
  }
}

// KCS[n][a][b] = sum_c B[c][a] B[c][b],  B[c][a] = sum_j joints[n][j][c] * C[j][a]   (j < NJ14)
#define KJ 14
#define KB 13
__global__ void __launch_bounds__(192) k_kcs_fwd(int N, int K, const float *__restrict__ joints, const float *__restrict__ Cm,
                                                 float *__restrict__ kcs) {
  __shared__ float sB[3][KB];
  __shared__ float sJ[KJ][3];
  int n = blockIdx.x, t = threadIdx.x;
  if (t < KJ * 3) sJ[t / 3][t % 3] = joints[((size_t)n * K + t / 3) * 3 + t % 3];
  __syncthreads();
  if (t < 3 * KB) {
    int c = t / KB, a = t % KB;
    float acc = 0.f;
    for (int j = 0; j < KJ; ++j) acc = fmaf(sJ[j][c], Cm[j * KB + a], acc);
    sB[c][a] = acc;
  }
  __syncthreads();
  if (t < KB * KB) {
    int a = t / KB, b = t % KB;
    kcs[(size_t)n * KB * KB + t] = sB[0][a] * sB[0][b] + sB[1][a] * sB[1][b] + sB[2][a] * sB[2][b];
  }
}

// d_joints[n][j][c] = sum_a C[j][a] dB[c][a],  dB[c][a] = sum_b (dK[a][b] + dK[b][a]) B[c][b];  joints >= 14 get 0.
__global__ void __launch_bounds__(192) k_kcs_bwd(int N, int K, const float *__restrict__ joints, const float *__restrict__ Cm,
                                                 const float *__restrict__ dK, float *__restrict__ d_joints) {
  __shared__ float sB[3][KB], sdB[3][KB];
  __shared__ float sJ[KJ][3];
  __shared__ float sK[KB][KB];
  int n = blockIdx.x, t = threadIdx.x;
  if (t < KJ * 3) sJ[t / 3][t % 3] = joints[((size_t)n * K + t / 3) * 3 + t % 3];
  if (t < KB * KB) sK[t / KB][t % KB] = dK[(size_t)n * KB * KB + t];
  __syncthreads();
  if (t < 3 * KB) {
    int c = t / KB, a = t % KB;
    float acc = 0.f;
    for (int j = 0; j < KJ; ++j) acc = fmaf(sJ[j][c], Cm[j * KB + a], acc);
    sB[c][a] = acc;
  }
  __syncthreads();
  if (t < 3 * KB) {
    int c = t / KB, a = t % KB;
    float acc = 0.f;
    for (int b = 0; b < KB; ++b) acc = fmaf(sK[a][b] + sK[b][a], sB[c][b], acc);
    sdB[c][a] = acc;
  }
  __syncthreads();
  for (int i = t; i < K * 3; i += blockDim.x) {
    int j = i / 3, c = i % 3;
    float acc = 0.f;
    if (j < KJ)
      for (int a = 0; a < KB; ++a) acc = fmaf(Cm[j * KB + a], sdB[c][a], acc);
    d_joints[((size_t)n * K + j) * 3 + c] = acc;
  }
}

// out = fake + alpha[n] * (real - fake), row-wise alpha (src/trainer.py:551-557)
__global__ void k_interp(size_t total, int row, const float *__restrict__ fake, const float *__restrict__ real,
                         const float *__restrict__ alpha, float *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float a = alpha[i / row];
  out[i] = fake[i] + a * (real[i] - fake[i]);
}

int launch_silhouette_csr(smplb_ctx *c, int B, int H, int W, const float *seg, float *points, int cap, int *offsets,
                          int *counts_scratch) {
  LAUNCH(c, "sil_count", B, 256, 0, k_sil_count, H * W, seg, counts_scratch);
  LAUNCH(c, "sil_scan", 1, 1024, 0, k_sil_scan, B, counts_scratch, offsets);
  LAUNCH(c, "sil_fill", B, 256, 0, k_sil_fill, H, W, seg, offsets, cap, points);
  return 0;
}

int launch_kcs(smplb_ctx *c, int N, int K, const float *joints, const float *Cm, float *kcs) {
  LAUNCH(c, "kcs_fwd", N, 192, 0, k_kcs_fwd, N, K, joints, Cm, kcs);
  return 0;
}

int launch_kcs_bwd(smplb_ctx *c, int N, int K, const float *joints, const float *Cm, const float *dK, float *d_joints) {
  LAUNCH(c, "kcs_bwd", N, 192, 0, k_kcs_bwd, N, K, joints, Cm, dK, d_joints);
  return 0;
}

int launch_interp(smplb_ctx *c, size_t total, int row, const float *fake, const float *real, const float *alpha, float *out) {
  LAUNCH(c, "interp", (unsigned)((total + 255) / 256), 256, 0, k_interp, total, row, fake, real, alpha, out);
  return 0;
}
