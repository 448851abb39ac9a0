// SURVEY.md section 8f ("next" rows adjacent to the hot path):
//   * silhouette extraction: tf.cast(tf.where(seg > 0)[:, :3], float32) of src/trainer.py:291,443
//     as an ordered on-device compaction straight into the CSR form the mesh loss consumes;
//   * get_kcs (src/models.py:123-139): KCS = B^T B with B = joints[:, :14]^T C, forward and
//     backward.  The reference builds an N x 13 x 13 x N intermediate to take its diagonal;
//     here it is 169 dot products of length 3 per sample.
#include "smplb_internal.h"

#define FULL 0xffffffffu

// counts[i] = #{seg[i] > 0}; bits (may be NULL): the flags as a bitmap, bit k of image i = seg[i][k] > 0 (words per
// image = ceil(HW / 32)) -- the fill reads 6 MB of flags instead of the 205 MB of masks a second time.
__global__ void __launch_bounds__(256) k_sil_count(int HW, const float *__restrict__ seg, int *__restrict__ counts,
                                                   unsigned *__restrict__ bits) {
  __shared__ int red[8];
  const int i = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  const float *s = seg + (size_t)i * HW;
  const int words = (HW + 31) / 32;
  const bool vec = (reinterpret_cast<uintptr_t>(s) & 15) == 0;
  int c = 0;
  // a warp walks 128 pixels (one float4 per lane) at a time; 8 lanes' nibbles make one 32-bit word
  for (int k0 = 128 * w; k0 < HW; k0 += 128 * 8) {
    const int k = k0 + 4 * lane;
    unsigned f = 0;
    if (vec && k + 4 <= HW) {
      const float4 v = *reinterpret_cast<const float4 *>(s + k);
      f = (unsigned)(v.x > 0.0f) | (unsigned)(v.y > 0.0f) << 1 | (unsigned)(v.z > 0.0f) << 2 | (unsigned)(v.w > 0.0f) << 3;
    } else {
      for (int j = 0; j < 4; ++j)
        if (k + j < HW) f |= (unsigned)(s[k + j] > 0.0f) << j;
    }
    c += __popc(f);
    unsigned word = f << (4 * (lane & 7));
    word |= __shfl_xor_sync(FULL, word, 1);
    word |= __shfl_xor_sync(FULL, word, 2);
    word |= __shfl_xor_sync(FULL, word, 4);
    const int wi = (k0 >> 5) + (lane >> 3);
    if (bits && (lane & 7) == 0 && wi < words) bits[(size_t)i * words + wi] = word;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
  if (lane == 0) red[w] = c;
  __syncthreads();
  if (t == 0) {
    int tot = 0;
    for (int q = 0; q < 8; ++q) tot += red[q];
    counts[i] = tot;
  }
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

// points[offsets[i] + rank] = (x = col, y = row) of the rank-th pixel > 0 in row-major order, from the bitmap k_sil_count
// wrote.  A warp owns a contiguous run of words and takes them one at a time, LANE = BIT: the lanes whose bit is set
// write consecutive points, so every store instruction covers one contiguous run of up to 256 bytes (with a run of
// words per thread the 32 lanes of a store wrote to 32 different places: 0.11 ms at B = 1024 whether the flags came from
// the masks or from the bitmap).
__global__ void __launch_bounds__(256) k_sil_fill_bits(int H, int W, const unsigned *__restrict__ bits,
                                                       const int *__restrict__ offsets, int cap, float *__restrict__ points) {
  __shared__ int wsum[8];
  const int i = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int HW = H * W, words = (HW + 31) / 32;
  const unsigned *bw = bits + (size_t)i * words;
  const int per = (words + 7) / 8;
  const int w0 = min(w * per, words), w1 = min(w0 + per, words);
  int c = 0;
  for (int k = w0 + lane; k < w1; k += 32) c += __popc(bw[k]);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
  if (lane == 0) wsum[w] = c;
  __syncthreads();
  if (c == 0) return;
  int pos = offsets[i];
  for (int q = 0; q < w; ++q) pos += wsum[q];
  const bool p8 = (reinterpret_cast<uintptr_t>(points) & 7) == 0;
  for (int k0 = w0; k0 < w1; k0 += 32) {
    const unsigned mine = k0 + lane < w1 ? bw[k0 + lane] : 0u;     // 32 words per load, handed out by shuffle
    const int nw = min(32, w1 - k0);
    for (int j = 0; j < nw; ++j) {
      const unsigned word = __shfl_sync(FULL, mine, j);
      if (word >> lane & 1u) {
        const int p = pos + __popc(word & ((1u << lane) - 1u));
        const int k = 32 * (k0 + j) + lane;
        const int row = k / W, col = k - row * W;
        if (p < cap) {
          // x = column (ops.py:123: silhouette_gt[:, 2]), y = row (ops.py:124: silhouette_gt[:, 1])
          if (p8) {
            *reinterpret_cast<float2 *>(points + 2 * (size_t)p) = make_float2((float)col, (float)row);
          } else {
            points[2 * (size_t)p + 0] = (float)col;
            points[2 * (size_t)p + 1] = (float)row;
          }
        }
      }
      pos += __popc(word);
    }
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

// ---- the critic's inputs and its gradient penalty, fused -------------------------
// (1) k_critic_inputs: everything src/trainer.py:548-557 builds for the interpolated forward pass in one launch:
//     x_hat = fake + alpha * (real - fake) with ELEMENT-wise alpha (tf.random.uniform(x.shape)) for the 3-D joints
//     [N,14,3], shapes [N,10] and rotations [N,23,3,3], and get_kcs(x_hat_joints) (src/models.py:123-139; the
//     reference forms an N x 13 x 13 x N intermediate to take its diagonal).  One CTA per sample.
__global__ void __launch_bounds__(256) k_critic_inputs(int K, const float *__restrict__ fj, const float *__restrict__ rj,
                                                       const float *__restrict__ aj, const float *__restrict__ fs,
                                                       const float *__restrict__ rs, const float *__restrict__ as_,
                                                       const float *__restrict__ fR, const float *__restrict__ rR,
                                                       const float *__restrict__ aR, const float *__restrict__ Cm,
                                                       float *__restrict__ oj, float *__restrict__ okcs,
                                                       float *__restrict__ os, float *__restrict__ oR) {
  __shared__ float sB[3][KB];
  __shared__ float sJ[KJ][3];
  const int n = blockIdx.x, t = threadIdx.x;
  for (int i = t; i < K * 3; i += 256) {
    const size_t g = (size_t)n * K * 3 + i;
    const float v = fj[g] + aj[g] * (rj[g] - fj[g]);
    oj[g] = v;
    if (i < KJ * 3) sJ[i / 3][i % 3] = v;
  }
  for (int i = t; i < 10 + 207; i += 256) {
    if (i < 10) {
      const size_t g = (size_t)n * 10 + i;
      os[g] = fs[g] + as_[g] * (rs[g] - fs[g]);
    } else {
      const size_t g = (size_t)n * 207 + (i - 10);
      oR[g] = fR[g] + aR[g] * (rR[g] - fR[g]);
    }
  }
  __syncthreads();
  if (t < 3 * KB) {
    const int c = t / KB, a = t % KB;
    float acc = 0.f;
    for (int j = 0; j < KJ; ++j) acc = fmaf(sJ[j][c], Cm[j * KB + a], acc);
    sB[c][a] = acc;
  }
  __syncthreads();
  if (t < KB * KB) {
    const int a = t / KB, b = t % KB;
    okcs[(size_t)n * KB * KB + t] = sB[0][a] * sB[0][b] + sB[1][a] * sB[1][b] + sB[2][a] * sB[2][b];
  }
}

// (2) k_critic_gp: what tf.gradients(out, [kcs, joints, shapes, Rs]) + compute_gradient_penalty (src/trainer.py:566-572,
//     src/ops.py:153-172) amount to once the critic has returned its partial derivatives: the gradient w.r.t. the
//     joints is the direct part plus the part through get_kcs (its backward, formed here per row), and the penalty
//     needs the column sums of all four gradients.  One pass over the inputs: a CTA walks 64 rows, per row forms
//     d_joints_total (optionally stored) and adds every column into its partial sums; the last CTA to finish (ticket)
//     adds the partials in chunk order (deterministic) and forms the penalty.
#define CG_ROWS 64
__global__ void __launch_bounds__(448) k_critic_gp(int M, int K, long long M_total, const float *__restrict__ joints,
                                                   const float *__restrict__ Cm, const float *__restrict__ g_kcs,
                                                   const float *__restrict__ g_j, const float *__restrict__ g_s,
                                                   const float *__restrict__ g_R, float *__restrict__ gj_total,
                                                   float *__restrict__ part, unsigned int *__restrict__ ticket,
                                                   float *__restrict__ col_sums, float *__restrict__ penalty) {
  __shared__ float sK[KB][KB], sJ[KJ][3], sB[3][KB], sdB[3][KB], sC[KJ * KB];
  __shared__ float sq[448], norms[4];
  __shared__ bool last;
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * CG_ROWS, m1 = min(M, m0 + CG_ROWS);
  if (t < KJ * KB) sC[t] = Cm[t];
  float acc = 0.f;   // this thread's column of the 428
  for (int m = m0; m < m1; ++m) {
    __syncthreads();
    if (t < KB * KB) {
      const float v = g_kcs[(size_t)m * KB * KB + t];
      sK[t / KB][t % KB] = v;
      acc += v;
    } else if (t >= 211 && t < 221) {
      acc += g_s[(size_t)m * 10 + (t - 211)];
    } else if (t >= 221 && t < 428) {
      acc += g_R[(size_t)m * 207 + (t - 221)];
    } else if (t >= 169 && t < 169 + KJ * 3) {
      const int i = t - 169;
      sJ[i / 3][i % 3] = joints[((size_t)m * K + i / 3) * 3 + i % 3];
    }
    __syncthreads();
    if (t < 3 * KB) {
      const int c = t / KB, a = t % KB;
      float b = 0.f;
      for (int j = 0; j < KJ; ++j) b = fmaf(sJ[j][c], sC[j * KB + a], b);
      sB[c][a] = b;
    }
    __syncthreads();
    if (t < 3 * KB) {
      const int c = t / KB, a = t % KB;
      float d = 0.f;
      for (int b = 0; b < KB; ++b) d = fmaf(sK[a][b] + sK[b][a], sB[c][b], d);
      sdB[c][a] = d;
    }
    __syncthreads();
    if (t >= 169 && t < 169 + KJ * 3) {
      const int i = t - 169, j = i / 3, c = i % 3;
      float d = g_j[(size_t)m * KJ * 3 + i];
      for (int a = 0; a < KB; ++a) d = fmaf(sC[j * KB + a], sdB[c][a], d);
      if (gj_total) gj_total[(size_t)m * KJ * 3 + i] = d;
      acc += d;
    }
  }
  if (t < SMPLB_GP_FLOATS) part[(size_t)blockIdx.x * SMPLB_GP_FLOATS + t] = acc;
  __threadfence();
  __syncthreads();
  if (t == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  float s = 0.f;
  if (t < SMPLB_GP_FLOATS) {
    for (int c = 0; c < (int)gridDim.x; ++c) s += __ldcg(part + (size_t)c * SMPLB_GP_FLOATS + t);
    col_sums[t] = s;
    const float mean = s / (float)M_total;
    sq[t] = mean * mean;
  }
  __syncthreads();
  if (t < 4) {
    const int beg[5] = {0, 169, 211, 221, 428};
    float q = 0.f;
    for (int i = beg[t]; i < beg[t + 1]; ++i) q += sq[i];
    norms[t] = sqrtf(q);
  }
  __syncthreads();
  if (t == 0) {
    float p = 0.f;
    for (int q = 0; q < 4; ++q) p += (1.0f - norms[q]) * (1.0f - norms[q]);
    *penalty = p;
    *ticket = 0;   // ready for the next call on this stream
  }
}

int launch_critic_inputs(smplb_ctx *c, int N, int K, const float *fj, const float *rj, const float *aj, const float *fs,
                         const float *rs, const float *as_, const float *fR, const float *rR, const float *aR,
                         const float *Cm, float *oj, float *okcs, float *os, float *oR) {
  LAUNCH(c, "critic_inputs_interp_kcs", N, 256, 0, k_critic_inputs, K, fj, rj, aj, fs, rs, as_, fR, rR, aR, Cm, oj, okcs, os, oR);
  return 0;
}

int launch_critic_gp(smplb_ctx *c, int M, int K, long long M_total, const float *joints, const float *Cm, const float *g_kcs,
                     const float *g_j, const float *g_s, const float *g_R, float *gj_total, float *part,
                     unsigned int *ticket, float *col_sums, float *penalty) {
  LAUNCH(c, "critic_gp_kcsbwd_colsum_penalty", cdiv(M, CG_ROWS), 448, 0, k_critic_gp, M, K, M_total, joints, Cm, g_kcs, g_j, g_s,
         g_R, gj_total, part, ticket, col_sums, penalty);
  return 0;
}

// the flags of B images as a bitmap in the context's workspace (grown on demand)
static int ensure_segbits(smplb_ctx *c, int B, int H, int W) {
  const size_t need = (size_t)B * ((size_t)(H * W + 31) / 32);
  if (need > c->ws_segbits_cap) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->ws_segbits) CUDA_TRY(cudaFree(c->ws_segbits));
    c->ws_segbits = nullptr;
    CUDA_TRY(cudaMalloc((void **)&c->ws_segbits, need * 4));
    c->ws_segbits_cap = need;
  }
  return 0;
}

int launch_silhouette_csr(smplb_ctx *c, int B, int H, int W, const float *seg, float *points, int cap, int *offsets,
                          int *counts_scratch) {
  TRY(ensure_segbits(c, B, H, W));
  LAUNCH(c, "sil_count", B, 256, 0, k_sil_count, H * W, seg, counts_scratch, c->ws_segbits);
  LAUNCH(c, "sil_scan", 1, 1024, 0, k_sil_scan, B, counts_scratch, offsets);
  if (cap > 0) LAUNCH(c, "sil_fill", B, 256, 0, k_sil_fill_bits, H, W, c->ws_segbits, offsets, cap, points);
  return 0;
}

// the fill alone, once the offsets (and the total) are known: from the bitmap the last launch_silhouette_csr of the same
// masks left in the workspace
int launch_silhouette_fill(smplb_ctx *c, int B, int H, int W, const float *seg, float *points, int cap, const int *offsets) {
  (void)seg;
  LAUNCH(c, "sil_fill", B, 256, 0, k_sil_fill_bits, H, W, c->ws_segbits, offsets, cap, points);
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
