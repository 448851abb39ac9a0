// Linear blend skinning, keypoint regression, projection (forward and backward) on FP32 CUDA
// cores.  The per-vertex 4x4 transforms T = W . A (batch_smpl.py:139-144) are formed in
// registers and never written to memory.
//
//   LBS                      src/tf_smpl/batch_smpl.py:139-149
//   keypoint regression      src/tf_smpl/batch_smpl.py:152-155
//   batch_orth_proj_idrot    src/tf_smpl/projection.py:23-33
//   reproject_vertices       src/tf_smpl/projection.py:45-56
//   kp_reprojection_loss     src/ops.py:35-47 (per-body partial sums)
#include <cuda_bf16.h>

#include "smplb_internal.h"

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------
// Forward skinning.  CTA = 128 vertices x SK_SPB samples.  Each thread keeps its vertex's 24
// skinning weights in registers (one coalesced float4 x 6 read per CTA); the sample loop
// stages the 24 3x4 transforms of SK_CH samples in shared memory and reads them as
// broadcast float4.
#define SK_VT 128
#define SK_SPB 32
#define SK_CH 8
__global__ void __launch_bounds__(SK_VT) k_skin_fwd(int B, int V, int Vp, const float *__restrict__ W,
                                                    const float *__restrict__ A, const float *__restrict__ v_posed,
                                                    float *__restrict__ verts) {
  __shared__ __align__(16) float sA[SK_CH][NJ * 12];
  int v = blockIdx.x * SK_VT + threadIdx.x;
  bool vok = v < V;
  int vc = vok ? v : V - 1;
  float w[NJ];
  const float4 *wrow = reinterpret_cast<const float4 *>(W + (size_t)vc * NJ);
#pragma unroll
  for (int q = 0; q < NJ / 4; ++q) {
    float4 t = __ldg(wrow + q);
    w[4 * q + 0] = t.x;
    w[4 * q + 1] = t.y;
    w[4 * q + 2] = t.z;
    w[4 * q + 3] = t.w;
  }
  int s_begin = blockIdx.y * SK_SPB;
  int s_end = min(B, s_begin + SK_SPB);
  for (int s0 = s_begin; s0 < s_end; s0 += SK_CH) {
    int ns = min(SK_CH, s_end - s0);
    __syncthreads();
    for (int i = threadIdx.x; i < ns * NJ * 12; i += SK_VT) sA[0][i] = A[(size_t)s0 * NJ * 12 + i];
    __syncthreads();
    for (int s = 0; s < ns; ++s) {
      float T[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] = 0.0f;
      const float4 *a4 = reinterpret_cast<const float4 *>(sA[s]);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float4 r0 = a4[3 * j + 0], r1 = a4[3 * j + 1], r2 = a4[3 * j + 2];
        float wj = w[j];
        T[0] = fmaf(wj, r0.x, T[0]); T[1] = fmaf(wj, r0.y, T[1]); T[2] = fmaf(wj, r0.z, T[2]); T[3] = fmaf(wj, r0.w, T[3]);
        T[4] = fmaf(wj, r1.x, T[4]); T[5] = fmaf(wj, r1.y, T[5]); T[6] = fmaf(wj, r1.z, T[6]); T[7] = fmaf(wj, r1.w, T[7]);
        T[8] = fmaf(wj, r2.x, T[8]); T[9] = fmaf(wj, r2.y, T[9]); T[10] = fmaf(wj, r2.z, T[10]); T[11] = fmaf(wj, r2.w, T[11]);
      }
      size_t b = (size_t)(s0 + s);
      const float *p = v_posed + b * (3 * (size_t)Vp) + vc;   // planar: coordinate c at c * Vp + v
      float p0 = p[0], p1 = p[Vp], p2 = p[2 * (size_t)Vp];
      if (vok) {
        float *o = verts + (b * V + v) * 3;
        o[0] = fmaf(T[0], p0, fmaf(T[1], p1, fmaf(T[2], p2, T[3])));
        o[1] = fmaf(T[4], p0, fmaf(T[5], p1, fmaf(T[6], p2, T[7])));
        o[2] = fmaf(T[8], p0, fmaf(T[9], p1, fmaf(T[10], p2, T[11])));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Keypoint regression from the sparse rows of joint_regressor (CSR by keypoint, built at
// create time from the dense [V,K] matrix: any sparsity pattern, dense included), then
// projection and the per-body part of the keypoint loss.  CTA = one body, warp k = keypoint k.
__global__ void k_joints(int B, int V, int K, const int *__restrict__ off, const int *__restrict__ idx,
                         const float *__restrict__ val, const float *__restrict__ verts,
                         const float *__restrict__ cam, const float *__restrict__ kp_gt, float *__restrict__ joints,
                         float *__restrict__ kp_pred, float *__restrict__ dkp, float *__restrict__ part,
                         int *__restrict__ cnt) {
  __shared__ float s_l[MAXK];
  __shared__ int s_c[MAXK];
  int b = blockIdx.x;
  int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *vb = verts + (size_t)b * V * 3;
  float x = 0.f, y = 0.f, z = 0.f;
  for (int e = off[k] + lane; e < off[k + 1]; e += 32) {
    int vi = idx[e];
    float wv = val[e];
    x = fmaf(wv, vb[3 * vi + 0], x);
    y = fmaf(wv, vb[3 * vi + 1], y);
    z = fmaf(wv, vb[3 * vi + 2], z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(FULL, x, o);
    y += __shfl_xor_sync(FULL, y, o);
    z += __shfl_xor_sync(FULL, z, o);
  }
  if (lane == 0) {
    size_t bk = (size_t)b * K + k;
    joints[bk * 3 + 0] = x;
    joints[bk * 3 + 1] = y;
    joints[bk * 3 + 2] = z;
    float l = 0.f;
    int cn = 0;
    if (cam) {
      float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
      float px = s * (x + tx), py = s * (y + ty);   // projection.py:27-33: translate, then scale
      if (kp_pred) {
        kp_pred[bk * 2 + 0] = px;
        kp_pred[bk * 2 + 1] = py;
      }
      if (kp_gt) {
        float gx = kp_gt[bk * 3 + 0], gy = kp_gt[bk * 3 + 1], vis = kp_gt[bk * 3 + 2];
        float dx = px - gx, dy = py - gy;
        l = vis * fabsf(dx) + vis * fabsf(dy);   // weights multiply by the vis VALUE (ops.py:42-45)
        cn = (vis != 0.0f) ? 2 : 0;              // num_present counts vis != 0, broadcast over x,y
        if (dkp) {
          dkp[bk * 2 + 0] = vis * (float)((dx > 0.f) - (dx < 0.f));
          dkp[bk * 2 + 1] = vis * (float)((dy > 0.f) - (dy < 0.f));
        }
      }
    }
    s_l[k] = l;
    s_c[k] = cn;
  }
  __syncthreads();
  if (threadIdx.x == 0 && part) {
    float tl = 0.f;
    int tc = 0;
    for (int q = 0; q < K; ++q) {
      tl += s_l[q];
      tc += s_c[q];
    }
    part[b] = tl;
    cnt[b] = tc;
  }
}

// ------------------------------------------------------------------------------------------
// Backward skinning.  CTA = SB_ST samples x one vertex range (blockIdx.y of VSPLIT).
//   phase 1 (thread = (4 vertices, sample)): g = d_verts + joint_regressor . d_joints,
//            T_R = sum_j W_vj A_R_j, dp = T_R^T g -> global; (g, [p;1]) -> shared.  Four vertices per thread so that
//            one A_R_j (three 16-byte shared-memory loads) feeds 36 FMAs instead of 9 (0.47 -> 0.435 ms at B = 1024;
//            the same treatment of phase 2 -- three rows per thread -- costs more in registers than it saves: 0.56 ms)
//   phase 2 (thread = (sample, 4 joints, row r)): dA[j][r][:] += W_vj * g_r * [p;1]
// Partial dA per vertex range is written to dA_part[split][b][24*12]; k_pose_bwd adds the
// VSPLIT partials in fixed order (deterministic, no float atomics).
#define SB_ST 16
#define SB_VT 64
#define SB_THREADS 288
struct SkinBwdSmem {
  float AR[SB_ST][NJ][12];   // A_R_j row-major, padded to three float4
  float W[NJ][SB_VT + 1];
  float GP[SB_VT][SB_ST][8];
  float DJ[SB_ST][MAXK][3];
};

__global__ void __launch_bounds__(SB_THREADS)
    k_skin_bwd(int B, int V, int Vreal, int K, int Vp_vp, int Vp_dp, const int *__restrict__ vmap,
               const float *__restrict__ W, const float *__restrict__ A, const float *__restrict__ v_posed,
               const float *__restrict__ d_verts, const float *__restrict__ d_joints, const int *__restrict__ voff,
               const int *__restrict__ vk, const float *__restrict__ vval, float *__restrict__ dp,
               float *__restrict__ dA_part, __nv_bfloat16 *__restrict__ dp16) {
  // V counts the vertices this launch walks: all of them, or (vmap != NULL) only the rows of
  // joint_regressor with a non-zero entry -- when no d_verts is given every other vertex has
  // g == 0 and contributes nothing.  W, voff are indexed by the walked index, v_posed /
  // d_verts by the real vertex vmap[v], dp by the walked index (compact when vmap != NULL).
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SkinBwdSmem &S = *reinterpret_cast<SkinBwdSmem *>(smem_raw);
  int tid = threadIdx.x;
  int s0 = blockIdx.x * SB_ST;
  int ns = min(SB_ST, B - s0);
  int tiles_total = cdiv_dev(V, SB_VT);
  int tiles_per = (tiles_total + (int)gridDim.y - 1) / (int)gridDim.y;
  int t_begin = blockIdx.y * tiles_per;
  int t_end = min(tiles_total, t_begin + tiles_per);

  for (int i = tid; i < SB_ST * NJ * 12; i += SB_THREADS) {
    int sl = i / (NJ * 12), r = i % (NJ * 12);
    int j = r / 12, e = r % 12;
    S.AR[sl][j][e] = (sl < ns && e < 9) ? A[((size_t)(s0 + sl) * NJ + j) * 12 + (e / 3) * 4 + (e % 3)] : 0.0f;
  }
  for (int i = tid; i < SB_ST * MAXK * 3; i += SB_THREADS) {
    int sl = i / (MAXK * 3), r = i % (MAXK * 3);
    int k = r / 3, cc = r % 3;
    S.DJ[sl][k][cc] = (d_joints && sl < ns && k < K) ? d_joints[((size_t)(s0 + sl) * K + k) * 3 + cc] : 0.0f;
  }
  // phase-2 role
  int p2_sl = tid / 18, p2_r18 = tid % 18;
  int p2_jg = p2_r18 / 3, p2_r = p2_r18 % 3;
  float acc[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int d = 0; d < 4; ++d) acc[q][d] = 0.0f;

  for (int t = t_begin; t < t_end; ++t) {
    int vbase = t * SB_VT;
    __syncthreads();  // previous tile's phase 2 done (also orders the AR/DJ fill on the first tile)
    for (int i = tid; i < SB_VT * NJ; i += SB_THREADS) {
      int vl = i / NJ, j = i % NJ;
      int v = vbase + vl;
      S.W[j][vl] = v < V ? W[(size_t)v * NJ + j] : 0.0f;
    }
    __syncthreads();
    if (tid < (SB_VT / 4) * SB_ST) {
      // vertices vg, vg + 16, vg + 32, vg + 48 of the tile for sample sl: T_R of all four from one pass over the joints
      const int vg = tid % (SB_VT / 4), sl = tid / (SB_VT / 4);
      float TR[4][9];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 9; ++e) TR[q][e] = 0.0f;
#pragma unroll 2
      for (int j = 0; j < NJ; ++j) {
        const float4 *ar4 = reinterpret_cast<const float4 *>(S.AR[sl][j]);
        const float4 a0 = ar4[0], a1 = ar4[1], a2 = ar4[2];
        const float ar[9] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float wj = S.W[j][vg + 16 * q];
#pragma unroll
          for (int e = 0; e < 9; ++e) TR[q][e] = fmaf(wj, ar[e], TR[q][e]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int vl = vg + 16 * q;
        const int v = vbase + vl;
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, p0 = 0.f, p1 = 0.f, p2 = 0.f, one = 0.f;
        if (v < V && sl < ns) {
          size_t b = (size_t)(s0 + sl);
          int vr = vmap ? vmap[v] : v;
          if (d_verts) {
            const float *dv = d_verts + (b * Vreal + vr) * 3;
            g0 = dv[0];
            g1 = dv[1];
            g2 = dv[2];
          }
          for (int e = voff[v]; e < voff[v + 1]; ++e) {
            int k = vk[e];
            float wv = vval[e];
            g0 = fmaf(wv, S.DJ[sl][k][0], g0);
            g1 = fmaf(wv, S.DJ[sl][k][1], g1);
            g2 = fmaf(wv, S.DJ[sl][k][2], g2);
          }
          const float *pp = v_posed + b * (3 * (size_t)Vp_vp) + vr;   // planar layouts
          p0 = pp[0];
          p1 = pp[Vp_vp];
          p2 = pp[2 * (size_t)Vp_vp];
          one = 1.0f;
          const float d0 = TR[q][0] * g0 + TR[q][3] * g1 + TR[q][6] * g2;   // dp = T_R^T g
          const float d1 = TR[q][1] * g0 + TR[q][4] * g1 + TR[q][7] * g2;
          const float d2 = TR[q][2] * g0 + TR[q][5] * g1 + TR[q][8] * g2;
          if (dp16) {
            // operand row of the tcgen05 blend-transpose GEMM: [hi | lo | hi], each 3 * Vp_dp wide, bf16 (16
            // significand bits between hi and lo, fp32's exponent range: no scaling needed)
            const size_t P3 = 3 * (size_t)Vp_dp;
            __nv_bfloat16 *o = dp16 + b * (3 * P3) + v;
            const float dd[3] = {d0, d1, d2};
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
              const __nv_bfloat16 hi = __float2bfloat16_rn(dd[cc]);
              const __nv_bfloat16 lo = __float2bfloat16_rn(dd[cc] - __bfloat162float(hi));
              o[cc * (size_t)Vp_dp] = hi;
              o[P3 + cc * (size_t)Vp_dp] = lo;
              o[2 * P3 + cc * (size_t)Vp_dp] = hi;
            }
          } else {
            float *o = dp + b * (3 * (size_t)Vp_dp) + v;
            o[0] = d0;
            o[Vp_dp] = d1;
            o[2 * (size_t)Vp_dp] = d2;
          }
        }
        float4 *dst = reinterpret_cast<float4 *>(S.GP[vl][sl]);
        dst[0] = make_float4(g0, g1, g2, 0.f);
        dst[1] = make_float4(p0, p1, p2, one);
      }
    }
    __syncthreads();
    if (p2_sl < SB_ST) {
#pragma unroll 4
      for (int vl = 0; vl < SB_VT; ++vl) {
        float gr = S.GP[vl][p2_sl][p2_r];
        float4 ph = *reinterpret_cast<const float4 *>(&S.GP[vl][p2_sl][4]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float wg = S.W[4 * p2_jg + q][vl] * gr;
          acc[q][0] = fmaf(wg, ph.x, acc[q][0]);
          acc[q][1] = fmaf(wg, ph.y, acc[q][1]);
          acc[q][2] = fmaf(wg, ph.z, acc[q][2]);
          acc[q][3] = fmaf(wg, ph.w, acc[q][3]);
        }
      }
    }
  }
  if (p2_sl < ns) {
    float *o = dA_part + ((size_t)blockIdx.y * B + (s0 + p2_sl)) * (NJ * 12);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int d = 0; d < 4; ++d) o[(4 * p2_jg + q) * 12 + 4 * p2_r + d] = acc[q][d];
  }
}

// ------------------------------------------------------------------------------------------
// Projection forward: out = s (X_xy + t); pixel: ((out + 1) * 0.5) * im_size.
__global__ void k_proj(size_t total, int N, const float *__restrict__ X, const float *__restrict__ cam, int pixel,
                       float im_w, float im_h, float *__restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  size_t b = i / N;
  float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
  float px = s * (X[i * 3 + 0] + tx), py = s * (X[i * 3 + 1] + ty);
  if (pixel) {
    px = ((px + 1.0f) * 0.5f) * im_w;
    py = ((py + 1.0f) * 0.5f) * im_h;
  }
  out[i * 2 + 0] = px;
  out[i * 2 + 1] = py;
}

// Projection backward.  CTA = one body; d_X = s g (z = 0), d_cam = (sum g.(xy+t), s sum g).
// g = d_out * gscale [/ *den] [* 0.5 im_size].  Fixed-order block reduction (deterministic).
// cnt != NULL: d_out is the mesh loss's raw gradient and (d_out + cnt) / denom is what k_mesh_grad_finish would have
// left there (the same operations, so the same bits) -- the step skips that pass over B x V x 2.
__global__ void __launch_bounds__(256) k_proj_bwd(int N, const float *__restrict__ X, const float *__restrict__ cam,
                                                  const float *__restrict__ d_out, int pixel, float im_w, float im_h,
                                                  float gscale, const long long *__restrict__ den, int accumulate_cam,
                                                  float *__restrict__ d_X, float *__restrict__ d_cam,
                                                  const int *__restrict__ cnt, float denom) {
  __shared__ float red[3][256];
  int b = blockIdx.x;
  float sc = gscale;
  if (den) {
    long long dv = *den;
    sc = dv > 0 ? gscale / (float)dv : 0.0f;
  }
  float fx = sc, fy = sc;
  if (pixel) {
    fx *= 0.5f * im_w;
    fy *= 0.5f * im_h;
  }
  float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
  float a_s = 0.f, a_x = 0.f, a_y = 0.f;
  for (int n = threadIdx.x; n < N; n += 256) {
    size_t i = (size_t)b * N + n;
    float ox = d_out[i * 2 + 0], oy = d_out[i * 2 + 1];
    if (cnt) {
      ox = (ox + (float)cnt[i * 2 + 0]) / denom;
      oy = (oy + (float)cnt[i * 2 + 1]) / denom;
    }
    float gx = ox * fx, gy = oy * fy;
    if (d_X) {
      d_X[i * 3 + 0] = s * gx;
      d_X[i * 3 + 1] = s * gy;
      d_X[i * 3 + 2] = 0.0f;
    }
    a_s += gx * (X[i * 3 + 0] + tx) + gy * (X[i * 3 + 1] + ty);
    a_x += gx;
    a_y += gy;
  }
  red[0][threadIdx.x] = a_s;
  red[1][threadIdx.x] = a_x;
  red[2][threadIdx.x] = a_y;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
      red[2][threadIdx.x] += red[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && d_cam) {
    float c0 = red[0][0], c1 = s * red[1][0], c2 = s * red[2][0];
    if (accumulate_cam) {
      c0 += d_cam[b * 3 + 0];
      c1 += d_cam[b * 3 + 1];
      c2 += d_cam[b * 3 + 2];
    }
    d_cam[b * 3 + 0] = c0;
    d_cam[b * 3 + 1] = c1;
    d_cam[b * 3 + 2] = c2;
  }
}

int launch_skin_fwd(smplb_ctx *c, int B, const float *A, const float *v_posed, float *verts) {
  dim3 grid(cdiv(c->V, SK_VT), cdiv(B, SK_SPB));
  LAUNCH(c, "skin_fwd", grid, SK_VT, 0, k_skin_fwd, B, c->V, c->Vp, c->d_W, A, v_posed, verts);
  return 0;
}

// act: verts is the compact [B][n_act][3] tensor and the CSR indexes active slots.
int launch_joints(smplb_ctx *c, int B, const float *verts, const float *cam, const float *kp_gt, float *joints,
                  float *kp_pred, float *dkp, float *part, int *cnt, bool act) {
  LAUNCH(c, act ? "joints_proj_kploss_active" : "joints_proj_kploss", B, 32 * c->K, 0, k_joints, B,
         act ? c->n_act : c->V, c->K, c->d_kcsr_off, act ? c->d_kcsr_slot : c->d_kcsr_idx, c->d_kcsr_val, verts, cam, kp_gt,
         joints, kp_pred, dkp, part, cnt);
  return 0;
}

int launch_skin_bwd(smplb_ctx *c, int B, const float *A, const float *v_posed, const float *d_verts,
                    const float *d_joints, float *dp, float *dA_part, int mode, void *dp16) {
  // mode 0: every vertex.  mode 1: active vertices, v_posed gathered from the full tensor.
  // mode 2: active vertices, v_posed is the compact v_posed_act (no gather).
  if (!(c->attr_done & 1u)) {
    CUDA_TRY(cudaFuncSetAttribute(k_skin_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SkinBwdSmem)));
    c->attr_done |= 1u;
  }
  dim3 grid(cdiv(B, SB_ST), skin_bwd_splits(B));
  if (mode == 2) {
    LAUNCH(c, "skin_bwd_active", grid, SB_THREADS, sizeof(SkinBwdSmem), k_skin_bwd, B, c->n_act, c->n_act, c->K, c->Vpa,
           c->Vpa, (const int *)nullptr, c->d_act_W, A, v_posed, (const float *)nullptr, d_joints, c->d_acsr_off,
           c->d_acsr_k, c->d_acsr_val, dp, dA_part, (__nv_bfloat16 *)nullptr);
  } else if (mode == 1) {
    LAUNCH(c, "skin_bwd_active_gather", grid, SB_THREADS, sizeof(SkinBwdSmem), k_skin_bwd, B, c->n_act, c->V, c->K, c->Vp,
           c->Vpa, c->d_act_idx, c->d_act_W, A, v_posed, (const float *)nullptr, d_joints, c->d_acsr_off,
           c->d_acsr_k, c->d_acsr_val, dp, dA_part, (__nv_bfloat16 *)nullptr);
  } else {
    LAUNCH(c, "skin_bwd", grid, SB_THREADS, sizeof(SkinBwdSmem), k_skin_bwd, B, c->V, c->V, c->K, c->Vp, c->Vp,
           (const int *)nullptr, c->d_W, A, v_posed, d_verts, d_joints, c->d_vcsr_off, c->d_vcsr_k, c->d_vcsr_val, dp,
           dA_part, (__nv_bfloat16 *)dp16);
  }
  return 0;
}

int launch_proj(smplb_ctx *c, int B, int N, const float *X, const float *cam, int pixel, float im_w, float im_h,
                float *out) {
  size_t total = (size_t)B * N;
  LAUNCH(c, "proj", (unsigned)((total + 255) / 256), 256, 0, k_proj, total, N, X, cam, pixel, im_w, im_h, out);
  return 0;
}

int launch_proj_bwd(smplb_ctx *c, int B, int N, const float *X, const float *cam, const float *d_out, int pixel,
                    float im_w, float im_h, float gscale, const long long *den, int accumulate_cam, float *d_X,
                    float *d_cam, const int *cnt, float denom) {
  LAUNCH(c, "proj_bwd", B, 256, 0, k_proj_bwd, N, X, cam, d_out, pixel, im_w, im_h, gscale, den, accumulate_cam, d_X,
         d_cam, cnt, denom);
  return 0;
}
