// Folded keypoint path: joints and their backward without touching the vertices.
//
// joints_k = sum_v JR_vk verts_v with verts_v = sum_j W_vj (A_R_j p_v + A_t_j), p_v = D_v x
//          = sum_j [ A_R_j (G_kj x) + A_t_j c_kj ],
//   G_kj = sum_v JR_vk W_vj D_v   (3 x 218, model constant),   c_kj = sum_v JR_vk W_vj,
//   x = [pose_feature(207) ; beta ; 1]
// -- exact algebra for src/tf_smpl/batch_smpl.py:110-155 restricted to the keypoint output
// (SURVEY.md section 0.6 / appendix B).  Per batch this is one GEMM U = x G^T ([B,218]x[218,1368])
// and a per-body contraction with the 24 transforms; the backward is the mirror image:
//   dA_j += dj_k (x) [u_kj ; c_kj],   du_kj = A_R_j^T dj_k,   dx = du G   ([B,1368]x[1368,218]).
// Both GEMMs run on tcgen05 (k_gemm_tc.cu) with fp16 operands split hi/lo along K; du is scaled
// per body by a power of two (its magnitude follows the loss scale) and unscaled in k_pose_bwd.
#include <cuda_fp16.h>
#include <math.h>

#include "smplb_internal.h"

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------- create time
// G[(k*24+j)*3+c][xk] and cc[k*24+j] in fp64 accumulation over the vertices keypoint k touches.
__global__ void __launch_bounds__(224) k_fold_build(int Vp, int pitch, const int *__restrict__ koff,
                                                    const int *__restrict__ kidx, const float *__restrict__ kval,
                                                    const float *__restrict__ W, const float *__restrict__ Dext,
                                                    float *__restrict__ G, float *__restrict__ cc) {
  int k = blockIdx.x / NJ, j = blockIdx.x % NJ;
  int xk = threadIdx.x;   // 0..223
  double a0 = 0, a1 = 0, a2 = 0, ac = 0;
  for (int e = koff[k]; e < koff[k + 1]; ++e) {
    int v = kidx[e];
    double w = (double)kval[e] * (double)W[(size_t)v * NJ + j];
    const float *d = Dext + (size_t)xk * pitch + v;
    a0 += w * (double)d[0];
    a1 += w * (double)d[Vp];
    a2 += w * (double)d[2 * (size_t)Vp];
    ac += w;
  }
  size_t n = (size_t)(k * NJ + j) * 3;
  G[(n + 0) * KX + xk] = (float)a0;
  G[(n + 1) * KX + xk] = (float)a1;
  G[(n + 2) * KX + xk] = (float)a2;
  if (xk == 0) cc[k * NJ + j] = (float)ac;
}

// GEMM operands.  G16 [NUp][256] pairs with the blend operand row x16 that k_pose_fwd writes anyway
// (same K map as Dt16 in k_blend_tc.cu): columns 0..206 G_pose (single fp16: it multiplies the small
// pose_feature), 207.. the shape columns as hi | lo | hi against beta_hi | beta_hi | beta_lo and the
// template column as hi | mid | lo against 1 | 1 | 1, so the ~1 m terms keep fp32 accuracy.
// Gt16 [224][3*NUp]: Gt_hi | Gt_hi | Gt_lo (pairs with du_hi | du_lo | du_hi).  Values are scaled
// by `scale` (a power of two) first.
__global__ void k_fold_operands(int nu, int nup, int NB, float scale, const float *__restrict__ G,
                                __half *__restrict__ G16, __half *__restrict__ Gt16) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nup * KX) return;
  int n = i / KX, xk = i % KX;
  float g = n < nu ? G[(size_t)n * KX + xk] * scale : 0.f;
  __half hi = __float2half_rn(g);
  __half lo = __float2half_rn(g - __half2float(hi));
  __half *r1 = G16 + (size_t)n * 256;
  if (xk < NPF) {
    r1[xk] = hi;
  } else if (xk < NPF + NB) {
    int bi = xk - NPF;
    r1[NPF + bi] = hi;
    r1[NPF + 10 + bi] = lo;
    r1[NPF + 20 + bi] = hi;
  } else if (xk == NPF + NB) {
    float mid = __half2float(lo);
    r1[NPF + 30] = hi;
    r1[NPF + 31] = lo;
    r1[NPF + 32] = __float2half_rn(g - __half2float(hi) - mid);
    for (int bi = NB; bi < 10; ++bi) {          // unused shape slots
      r1[NPF + bi] = r1[NPF + 10 + bi] = r1[NPF + 20 + bi] = __float2half_rn(0.f);
    }
    for (int k = NPF + 33; k < 256; ++k) r1[k] = __float2half_rn(0.f);
  }
  __half *r2 = Gt16 + (size_t)xk * (3 * nup);
  r2[n] = hi;
  r2[nup + n] = hi;
  r2[2 * nup + n] = lo;
}

// ---------------------------------------------------------------------------------- forward
// CTA = one body, warp k = keypoint k, lane j = joint j: joints_k = sum_j A_j [u_kj ; c_kj],
// then projection and the per-body part of the keypoint loss (same tail as k_joints).
__global__ void k_fold_fwd(int B, int K, int ldu, const float *__restrict__ U, const float *__restrict__ cc,
                           const float *__restrict__ A, const float *__restrict__ cam, const float *__restrict__ kp_gt,
                           float *__restrict__ joints, float *__restrict__ kp_pred, float *__restrict__ dkp,
                           float *__restrict__ part, int *__restrict__ cnt) {
  __shared__ float s_l[MAXK];
  __shared__ int s_c[MAXK];
  int b = blockIdx.x;
  int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float x = 0.f, y = 0.f, z = 0.f;
  if (lane < NJ) {
    const float *u = U + (size_t)b * ldu + (size_t)(k * NJ + lane) * 3;
    float u0 = u[0], u1 = u[1], u2 = u[2], c = cc[k * NJ + lane];
    const float4 *a = reinterpret_cast<const float4 *>(A + ((size_t)b * NJ + lane) * 12);
    float4 r0 = a[0], r1 = a[1], r2 = a[2];
    x = fmaf(r0.x, u0, fmaf(r0.y, u1, fmaf(r0.z, u2, r0.w * c)));
    y = fmaf(r1.x, u0, fmaf(r1.y, u1, fmaf(r1.z, u2, r1.w * c)));
    z = fmaf(r2.x, u0, fmaf(r2.y, u1, fmaf(r2.z, u2, r2.w * c)));
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
      float px = s * (x + tx), py = s * (y + ty);
      if (kp_pred) {
        kp_pred[bk * 2 + 0] = px;
        kp_pred[bk * 2 + 1] = py;
      }
      if (kp_gt) {
        float gx = kp_gt[bk * 3 + 0], gy = kp_gt[bk * 3 + 1], vis = kp_gt[bk * 3 + 2];
        float dx = px - gx, dy = py - gy;
        l = vis * fabsf(dx) + vis * fabsf(dy);
        cn = (vis != 0.0f) ? 2 : 0;
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

// --------------------------------------------------------------------------------- backward
// CTA = one body, warp k, lane j.  du_kj = A_R_j^T dj_k -> fp16 hi/lo operand row scaled by
// 2^-e_b (e_b from the body's max |du|); dA_j = sum_k dj_k (x) [u_kj ; c_kj] summed over the
// warps in fixed order -> dA_part[0][b] (the other VSPLIT partials are not read: n_parts = 1).
__global__ void k_fold_bwd(int B, int K, int ldu, int nup, const float *__restrict__ U, const float *__restrict__ cc,
                           const float *__restrict__ A, const float *__restrict__ d_joints, float *__restrict__ dA,
                           __half *__restrict__ du16, float *__restrict__ rowscale) {
  __shared__ float s_dA[MAXK][NJ * 12];
  __shared__ float s_max[MAXK];
  int b = blockIdx.x;
  int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float du0 = 0.f, du1 = 0.f, du2 = 0.f;
  if (lane < NJ) {
    const float *dj = d_joints + ((size_t)b * K + k) * 3;
    float g0 = dj[0], g1 = dj[1], g2 = dj[2];
    const float *u = U + (size_t)b * ldu + (size_t)(k * NJ + lane) * 3;
    float u4[4] = {u[0], u[1], u[2], cc[k * NJ + lane]};
    const float *a = A + ((size_t)b * NJ + lane) * 12;
    du0 = a[0] * g0 + a[4] * g1 + a[8] * g2;     // A_R^T dj
    du1 = a[1] * g0 + a[5] * g1 + a[9] * g2;
    du2 = a[2] * g0 + a[6] * g1 + a[10] * g2;
    float gg[3] = {g0, g1, g2};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int d = 0; d < 4; ++d) s_dA[k][lane * 12 + 4 * r + d] = gg[r] * u4[d];
  }
  float m = fmaxf(fabsf(du0), fmaxf(fabsf(du1), fabsf(du2)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if (lane == 0) s_max[k] = m;
  __syncthreads();
  float mx = 0.f;
  for (int q = 0; q < K; ++q) mx = fmaxf(mx, s_max[q]);
  // power-of-two scale that brings max |du| to [1, 2): exact to apply and to undo
  int e = 0;
  if (mx > 0.f && isfinite(mx)) frexpf(mx, &e);
  float inv = ldexpf(1.0f, 1 - e), sc = ldexpf(1.0f, e - 1);
  if (!(mx > 0.f)) {
    inv = 1.0f;
    sc = 1.0f;
  }
  if (threadIdx.x == 0) rowscale[b] = sc;
  if (lane < NJ) {
    __half *row = du16 + (size_t)b * (3 * nup);
    int n = (k * NJ + lane) * 3;
    float d3[3] = {du0 * inv, du1 * inv, du2 * inv};
#pragma unroll
    for (int cI = 0; cI < 3; ++cI) {
      __half hi = __float2half_rn(d3[cI]);
      __half lo = __float2half_rn(d3[cI] - __half2float(hi));
      row[n + cI] = hi;
      row[nup + n + cI] = lo;
      row[2 * nup + n + cI] = hi;
    }
  }
  for (int i = threadIdx.x; i < NJ * 12; i += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < K; ++q) s += s_dA[q][i];
    dA[(size_t)b * (NJ * 12) + i] = s;
  }
}

// ------------------------------------------------------------------ warp-per-body variants
// One warp per body, 8 bodies per CTA: far fewer, fuller CTAs than the CTA-per-body kernels
// above (which remain as the readable reference of the same arithmetic).
#define FW_WARPS 8

// Forward: lane k owns keypoint k and loops over the 24 joints (A staged in shared memory).
__global__ void __launch_bounds__(32 * FW_WARPS)
    k_fold_fwd_w(int B, int K, int ldu, const float *__restrict__ U, const float *__restrict__ cc,
                 const float *__restrict__ A, const float *__restrict__ cam, const float *__restrict__ kp_gt,
                 float *__restrict__ joints, float *__restrict__ kp_pred, float *__restrict__ dkp,
                 float *__restrict__ part, int *__restrict__ cnt) {
  __shared__ __align__(16) float sA[FW_WARPS][NJ * 12];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x * FW_WARPS + w;
  if (b >= B) return;
  for (int i = lane; i < NJ * 12; i += 32) sA[w][i] = A[(size_t)b * NJ * 12 + i];
  __syncwarp();
  float x = 0.f, y = 0.f, z = 0.f, l = 0.f;
  int cn = 0;
  if (lane < K) {
    const float *u = U + (size_t)b * ldu + (size_t)lane * NJ * 3;
    const float *c = cc + lane * NJ;
#pragma unroll 4
    for (int j = 0; j < NJ; ++j) {
      float u0 = u[3 * j], u1 = u[3 * j + 1], u2 = u[3 * j + 2], cj = c[j];
      const float4 *a = reinterpret_cast<const float4 *>(sA[w] + j * 12);
      float4 r0 = a[0], r1 = a[1], r2 = a[2];
      x += fmaf(r0.x, u0, fmaf(r0.y, u1, fmaf(r0.z, u2, r0.w * cj)));
      y += fmaf(r1.x, u0, fmaf(r1.y, u1, fmaf(r1.z, u2, r1.w * cj)));
      z += fmaf(r2.x, u0, fmaf(r2.y, u1, fmaf(r2.z, u2, r2.w * cj)));
    }
    size_t bk = (size_t)b * K + lane;
    joints[bk * 3 + 0] = x;
    joints[bk * 3 + 1] = y;
    joints[bk * 3 + 2] = z;
    if (cam) {
      float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
      float px = s * (x + tx), py = s * (y + ty);
      if (kp_pred) {
        kp_pred[bk * 2 + 0] = px;
        kp_pred[bk * 2 + 1] = py;
      }
      if (kp_gt) {
        float gx = kp_gt[bk * 3 + 0], gy = kp_gt[bk * 3 + 1], vis = kp_gt[bk * 3 + 2];
        float dx = px - gx, dy = py - gy;
        l = vis * fabsf(dx) + vis * fabsf(dy);
        cn = (vis != 0.0f) ? 2 : 0;
        if (dkp) {
          dkp[bk * 2 + 0] = vis * (float)((dx > 0.f) - (dx < 0.f));
          dkp[bk * 2 + 1] = vis * (float)((dy > 0.f) - (dy < 0.f));
        }
      }
    }
  }
  if (part) {
    // fixed-order sum over the keypoints (lane 0 adds k = 0, 1, ... in turn)
    float tl = 0.f;
    int tc = 0;
    for (int q = 0; q < K; ++q) {
      tl += __shfl_sync(FULL, l, q);
      tc += __shfl_sync(FULL, cn, q);
    }
    if (lane == 0) {
      part[b] = tl;
      cnt[b] = tc;
    }
  }
}

// Backward: (phase A, lane k) dj_k either given or formed from the unscaled keypoint-loss
// gradient, the camera and gscale / *den -- then d_cam falls out too; (phase B, lane j)
// du_kj = A_R_j^T dj_k, dA_j = sum_k dj_k (x) [u_kj ; c_kj] with no cross-lane reduction.
__global__ void __launch_bounds__(32 * FW_WARPS)
    k_fold_bwd_w(int B, int K, int ldu, int nup, const float *__restrict__ U, const float *__restrict__ cc,
                 const float *__restrict__ A, const float *__restrict__ d_joints, const float *__restrict__ dkp,
                 const float *__restrict__ joints, const float *__restrict__ cam, float gscale,
                 const long long *__restrict__ den, float *__restrict__ d_cam, float *__restrict__ dA,
                 __half *__restrict__ du16, float *__restrict__ rowscale) {
  __shared__ float sdj[FW_WARPS][MAXK * 3];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x * FW_WARPS + w;
  if (b >= B) return;
  if (dkp) {
    float sc = gscale;
    if (den) {
      long long dv = *den;
      sc = dv > 0 ? gscale / (float)dv : 0.0f;
    }
    float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
    float gx = 0.f, gy = 0.f, as = 0.f;
    if (lane < K) {
      size_t bk = (size_t)b * K + lane;
      gx = dkp[bk * 2 + 0] * sc;
      gy = dkp[bk * 2 + 1] * sc;
      as = gx * (joints[bk * 3 + 0] + tx) + gy * (joints[bk * 3 + 1] + ty);
      sdj[w][lane * 3 + 0] = s * gx;
      sdj[w][lane * 3 + 1] = s * gy;
      sdj[w][lane * 3 + 2] = 0.0f;
    }
    if (d_cam) {
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      for (int q = 0; q < K; ++q) {   // fixed order
        c0 += __shfl_sync(FULL, as, q);
        c1 += __shfl_sync(FULL, gx, q);
        c2 += __shfl_sync(FULL, gy, q);
      }
      if (lane == 0) {
        d_cam[b * 3 + 0] = c0;
        d_cam[b * 3 + 1] = s * c1;
        d_cam[b * 3 + 2] = s * c2;
      }
    }
  } else {
    for (int i = lane; i < K * 3; i += 32) sdj[w][i] = d_joints[(size_t)b * K * 3 + i];
  }
  __syncwarp();
  int j = lane < NJ ? lane : NJ - 1;
  const float *a = A + ((size_t)b * NJ + j) * 12;
  float ar[9] = {a[0], a[1], a[2], a[4], a[5], a[6], a[8], a[9], a[10]};
  const float *u = U + (size_t)b * ldu + (size_t)j * 3;
  float acc[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) acc[e] = 0.f;
  float m = 0.f;
  for (int k = 0; k < K; ++k) {
    float g0 = sdj[w][3 * k], g1 = sdj[w][3 * k + 1], g2 = sdj[w][3 * k + 2];
    float du0 = ar[0] * g0 + ar[3] * g1 + ar[6] * g2;
    float du1 = ar[1] * g0 + ar[4] * g1 + ar[7] * g2;
    float du2 = ar[2] * g0 + ar[5] * g1 + ar[8] * g2;
    m = fmaxf(m, fmaxf(fabsf(du0), fmaxf(fabsf(du1), fabsf(du2))));
    const float *uk = u + (size_t)k * NJ * 3;
    float u4[4] = {uk[0], uk[1], uk[2], cc[k * NJ + j]};
    float gg[3] = {g0, g1, g2};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int d = 0; d < 4; ++d) acc[4 * r + d] = fmaf(gg[r], u4[d], acc[4 * r + d]);
  }
  if (lane >= NJ) m = 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  int e = 0;
  float inv = 1.0f, sc2 = 1.0f;
  if (m > 0.f && isfinite(m)) {
    frexpf(m, &e);
    inv = ldexpf(1.0f, 1 - e);
    sc2 = ldexpf(1.0f, e - 1);
  }
  if (lane == 0) rowscale[b] = sc2;
  if (lane < NJ) {
    float *o = dA + ((size_t)b * NJ + j) * 12;
#pragma unroll
    for (int q = 0; q < 12; ++q) o[q] = acc[q];
    __half *row = du16 + (size_t)b * (3 * nup);
    for (int k = 0; k < K; ++k) {
      float g0 = sdj[w][3 * k], g1 = sdj[w][3 * k + 1], g2 = sdj[w][3 * k + 2];
      float d3[3] = {(ar[0] * g0 + ar[3] * g1 + ar[6] * g2) * inv, (ar[1] * g0 + ar[4] * g1 + ar[7] * g2) * inv,
                     (ar[2] * g0 + ar[5] * g1 + ar[8] * g2) * inv};
      int n = (k * NJ + j) * 3;
#pragma unroll
      for (int cI = 0; cI < 3; ++cI) {
        __half hi = __float2half_rn(d3[cI]);
        __half lo = __float2half_rn(d3[cI] - __half2float(hi));
        row[n + cI] = hi;
        row[nup + n + cI] = lo;
        row[2 * nup + n + cI] = hi;
      }
    }
  }
}

// One generator step in one kernel (smplb_step with the keypoint loss only): forward (joints,
// projection, loss partials) and backward (du, dA, d_cam) of the folded keypoint path read U and A
// once.  The loss gradient is formed for a unit upstream scale -- the global 1 / num_present is
// not known before the batch-wide reduction (and, across GPUs, the all-reduce) -- and
// k_pose_bwd applies gscale / num_present to its outputs and to d_cam (everything in between is
// linear), so the reduction runs beside the backward GEMM instead of in front of it.
#define FS_WARPS 4
#define FS_MAXU 1536   // floats of one U row staged per warp (3 * 24 * K rounded up to 128; K <= 21)
__global__ void __launch_bounds__(32 * FS_WARPS)
    k_fold_step_w(int B, int K, int ldu, int nup, const float *__restrict__ U, const float *__restrict__ cc,
                  const float *__restrict__ A, const float *__restrict__ cam, const float *__restrict__ kp_gt,
                  float *__restrict__ joints, float *__restrict__ kp_pred, float *__restrict__ part,
                  int *__restrict__ cnt, float *__restrict__ d_cam, float *__restrict__ dA, __half *__restrict__ du16,
                  float *__restrict__ rowscale) {
  __shared__ __align__(16) float sU[FS_WARPS][FS_MAXU];
  __shared__ __align__(16) float sA[FS_WARPS][NJ * 12];
  __shared__ float sdj[FS_WARPS][MAXK * 3];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x * FS_WARPS + w;
  if (b >= B) return;
  // The kernel is a chain of dependent global loads if U is read where it is used (one latency per
  // joint / keypoint iteration): stage the body's U row (5.5 KB) and A with independent vector
  // loads first, then both phases run out of shared memory.
  {
    const float4 *src = reinterpret_cast<const float4 *>(U + (size_t)b * ldu);
    float4 *dst = reinterpret_cast<float4 *>(sU[w]);
    const int n4 = (K * NJ * 3 + 3) / 4;
    for (int i = lane; i < n4; i += 32) dst[i] = src[i];
    const float4 *asrc = reinterpret_cast<const float4 *>(A + (size_t)b * NJ * 12);
    float4 *adst = reinterpret_cast<float4 *>(sA[w]);
    for (int i = lane; i < NJ * 3; i += 32) adst[i] = asrc[i];
  }
  __syncwarp();
  const float *Ub = sU[w];
  const float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
  // ---- forward, lane k owns keypoint k (k_fold_fwd_w)
  float l = 0.f, gx = 0.f, gy = 0.f, as = 0.f;
  int cn = 0;
  if (lane < K) {
    const float *u = Ub + lane * NJ * 3;
    const float *c = cc + lane * NJ;
    float x = 0.f, y = 0.f, z = 0.f;
#pragma unroll 4
    for (int j = 0; j < NJ; ++j) {
      float u0 = u[3 * j], u1 = u[3 * j + 1], u2 = u[3 * j + 2], cj = c[j];
      const float4 *a = reinterpret_cast<const float4 *>(sA[w] + j * 12);
      float4 r0 = a[0], r1 = a[1], r2 = a[2];
      x += fmaf(r0.x, u0, fmaf(r0.y, u1, fmaf(r0.z, u2, r0.w * cj)));
      y += fmaf(r1.x, u0, fmaf(r1.y, u1, fmaf(r1.z, u2, r1.w * cj)));
      z += fmaf(r2.x, u0, fmaf(r2.y, u1, fmaf(r2.z, u2, r2.w * cj)));
    }
    size_t bk = (size_t)b * K + lane;
    joints[bk * 3 + 0] = x;
    joints[bk * 3 + 1] = y;
    joints[bk * 3 + 2] = z;
    float px = s * (x + tx), py = s * (y + ty);
    if (kp_pred) {
      kp_pred[bk * 2 + 0] = px;
      kp_pred[bk * 2 + 1] = py;
    }
    float gtx = kp_gt[bk * 3 + 0], gty = kp_gt[bk * 3 + 1], vis = kp_gt[bk * 3 + 2];
    float dx = px - gtx, dy = py - gty;
    l = vis * fabsf(dx) + vis * fabsf(dy);
    cn = (vis != 0.0f) ? 2 : 0;
    gx = vis * (float)((dx > 0.f) - (dx < 0.f));   // d loss / d kp_pred for a unit scale
    gy = vis * (float)((dy > 0.f) - (dy < 0.f));
    as = gx * (x + tx) + gy * (y + ty);
    sdj[w][lane * 3 + 0] = s * gx;
    sdj[w][lane * 3 + 1] = s * gy;
    sdj[w][lane * 3 + 2] = 0.0f;
  }
  {
    // fixed-order sums over the keypoints (every lane adds k = 0, 1, ... in turn)
    float tl = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    int tc = 0;
    for (int q = 0; q < K; ++q) {
      tl += __shfl_sync(FULL, l, q);
      tc += __shfl_sync(FULL, cn, q);
      c0 += __shfl_sync(FULL, as, q);
      c1 += __shfl_sync(FULL, gx, q);
      c2 += __shfl_sync(FULL, gy, q);
    }
    if (lane == 0) {
      part[b] = tl;
      cnt[b] = tc;
      d_cam[b * 3 + 0] = c0;          // unscaled; k_pose_bwd multiplies by gscale / num_present
      d_cam[b * 3 + 1] = s * c1;
      d_cam[b * 3 + 2] = s * c2;
    }
  }
  __syncwarp();
  // ---- backward, lane j owns joint j (k_fold_bwd_w): du_kj = A_R_j^T dj_k, dA_j = sum_k dj_k (x) [u_kj ; c_kj]
  int j = lane < NJ ? lane : NJ - 1;
  const float *a = sA[w] + j * 12;
  float ar[9] = {a[0], a[1], a[2], a[4], a[5], a[6], a[8], a[9], a[10]};
  const float *u = Ub + j * 3;
  float acc[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) acc[e] = 0.f;
  float m = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float g0 = sdj[w][3 * k], g1 = sdj[w][3 * k + 1], g2 = sdj[w][3 * k + 2];
    float du0 = ar[0] * g0 + ar[3] * g1 + ar[6] * g2;
    float du1 = ar[1] * g0 + ar[4] * g1 + ar[7] * g2;
    float du2 = ar[2] * g0 + ar[5] * g1 + ar[8] * g2;
    m = fmaxf(m, fmaxf(fabsf(du0), fmaxf(fabsf(du1), fabsf(du2))));
    const float *uk = u + k * NJ * 3;
    float u4[4] = {uk[0], uk[1], uk[2], cc[k * NJ + j]};
    if (lane < NJ) {
      // u_kj has been read: its three floats now stage du_kj (unit row scale) for the coalesced fp16 row below
      float *d = sU[w] + (k * NJ + j) * 3;
      d[0] = du0;
      d[1] = du1;
      d[2] = du2;
    }
    float gg[3] = {g0, g1, g2};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int d = 0; d < 4; ++d) acc[4 * r + d] = fmaf(gg[r], u4[d], acc[4 * r + d]);
  }
  if (lane >= NJ) m = 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  // power-of-two scale that brings max |du| to [1, 2): exact to apply and to undo
  int e = 0;
  float inv = 1.0f, sc2 = 1.0f;
  if (m > 0.f && isfinite(m)) {
    frexpf(m, &e);
    inv = ldexpf(1.0f, 1 - e);
    sc2 = ldexpf(1.0f, e - 1);
  }
  if (lane == 0) rowscale[b] = sc2;
  if (lane < NJ) {
    float *o = dA + ((size_t)b * NJ + j) * 12;
#pragma unroll
    for (int q = 0; q < 12; ++q) o[q] = acc[q];
  }
  // du16 row = hi | lo | hi (three blocks of nup halves) from the staged fp32 values, scaled by the row's power of
  // two (exact), with coalesced 4-byte stores instead of 2-byte stores 144 B apart
  __syncwarp();
  {
    __half2 *row = reinterpret_cast<__half2 *>(du16 + (size_t)b * (3 * nup));
    const int nu = K * NJ * 3, half_nup = nup / 2;
    for (int i = lane; i < half_nup; i += 32) {
      float v0 = 2 * i < nu ? sU[w][2 * i] * inv : 0.f, v1 = 2 * i + 1 < nu ? sU[w][2 * i + 1] * inv : 0.f;
      __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
      __half2 hi = __halves2half2(h0, h1);
      __half2 lo = __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
      row[i] = hi;
      row[half_nup + i] = lo;
      row[2 * half_nup + i] = hi;
    }
  }
}

// ------------------------------------------------------------------------------------- host

__global__ void k_absmax_f(size_t n, const float *__restrict__ x, unsigned int *__restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

int fold_init(smplb_ctx *c) {
  c->fold_ok = false;
  if (!c->tc_ok) return 0;
  int nu = 3 * NJ * c->K;
  int nup = cdiv(nu, 128) * 128;
  c->fold_nu = nu;
  c->fold_nup = nup;
  CUDA_TRY(cudaMalloc((void **)&c->d_G, (size_t)nu * KX * 4));
  CUDA_TRY(cudaMalloc((void **)&c->d_cc, (size_t)NJ * c->K * 4));
  k_fold_build<<<c->K * NJ, 224, 0, c->stream>>>(c->Vp, c->pitch, c->d_kcsr_off, c->d_kcsr_idx, c->d_kcsr_val, c->d_W,
                                                c->d_Dext, c->d_G, c->d_cc);
  unsigned int *d_max = nullptr;
  CUDA_TRY(cudaMalloc((void **)&d_max, 4));
  CUDA_TRY(cudaMemsetAsync(d_max, 0, 4, c->stream));
  k_absmax_f<<<148, 256, 0, c->stream>>>((size_t)nu * KX, c->d_G, d_max);
  unsigned int bits = 0;
  CUDA_TRY(cudaMemcpyAsync(&bits, d_max, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  cudaFree(d_max);
  float mx;
  memcpy(&mx, &bits, 4);
  int s = 0;
  if (mx > 0.f && isfinite(mx)) s = (int)floorf(log2f(1024.0f / mx));
  s = s < -8 ? -8 : (s > 24 ? 24 : s);
  c->fold_scale = ldexpf(1.0f, s);
  c->fold_inv_scale = ldexpf(1.0f, -s);
  CUDA_TRY(cudaMalloc((void **)&c->d_G16, (size_t)nup * 256 * 2));
  CUDA_TRY(cudaMalloc((void **)&c->d_Gt16, (size_t)KX * 3 * nup * 2));
  k_fold_operands<<<cdiv(nup * KX, 256), 256, 0, c->stream>>>(nu, nup, c->NB, c->fold_scale, c->d_G, (__half *)c->d_G16,
                                                             (__half *)c->d_Gt16);
  c->launches += 3;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  TRY(tc_make_map(c->map_g1, 0, c->d_G16, 256, (uint64_t)nup, 256 * 2, 64, 128));
  TRY(tc_make_map(c->map_g2, 0, c->d_Gt16, (uint64_t)3 * nup, (uint64_t)KX, (uint64_t)3 * nup * 2, 64, 128));
  c->fold_ok = true;
  return 0;
}

// U [B][nup] = x G^T, x16 = the blend operand rows [B][256].
int launch_fold_gemm_u(smplb_ctx *c, int B, const void *x16) {
  RET_IF(!c->fold_ok, SMPLB_ESTATE, "folded keypoint path is not initialised");
  return launch_gemm_tc(c, "fold_gemm_u", B, c->fold_nup, 256, x16, c->map_g1, c->ws_U, c->fold_nup, 1, c->fold_inv_scale);
}

// joints / projection / kp-loss partials from U and A.
int launch_fold_fwd(smplb_ctx *c, int B, const void *x16b, const float *A, const float *cam, const float *kp_gt,
                    float *joints, float *kp_pred, float *dkp, float *part, int *cnt) {
  RET_IF(!c->fold_ok, SMPLB_ESTATE, "folded keypoint path is not initialised");
  (void)x16b;   // U was produced by launch_fold_gemm_u
  if (c->fold_warp_kernels) {
    LAUNCH(c, "fold_joints_proj_kploss", cdiv(B, FW_WARPS), 32 * FW_WARPS, 0, k_fold_fwd_w, B, c->K, c->fold_nup, c->ws_U,
           c->d_cc, A, cam, kp_gt, joints, kp_pred, dkp, part, cnt);
  } else {
    LAUNCH(c, "fold_joints_proj_kploss_cta", B, 32 * c->K, 0, k_fold_fwd, B, c->K, c->fold_nup, c->ws_U, c->d_cc, A, cam,
           kp_gt, joints, kp_pred, dkp, part, cnt);
  }
  return 0;
}

// dA (one partial) and dx partials [ksplit][rows_per][KX] (+ per-body scale) from d_joints.
// Either d_joints is given, or (dkp, joints, cam, gscale, den) from which the kernel forms
// d_joints = s * dkp * gscale / *den itself and also writes d_cam.
int launch_fold_bwd(smplb_ctx *c, int B, const float *A, const float *d_joints, const float *dkp, const float *joints,
                    const float *cam, float gscale, const long long *den, float *d_cam, float *dA_part, float *dx_part,
                    int ksplit) {
  RET_IF(!c->fold_ok, SMPLB_ESTATE, "folded keypoint path is not initialised");
  if (c->fold_warp_kernels) {
    LAUNCH(c, "fold_bwd_du_dA", cdiv(B, FW_WARPS), 32 * FW_WARPS, 0, k_fold_bwd_w, B, c->K, c->fold_nup, c->fold_nup, c->ws_U,
           c->d_cc, A, d_joints, dkp, joints, cam, gscale, den, d_cam, dA_part, (__half *)c->ws_du16, c->ws_rowscale);
  } else {
    RET_IF(!d_joints, SMPLB_EINVAL, "the CTA-per-body fold backward needs d_joints");
    LAUNCH(c, "fold_bwd_du_dA_cta", B, 32 * c->K, 0, k_fold_bwd, B, c->K, c->fold_nup, c->fold_nup, c->ws_U, c->d_cc, A,
           d_joints, dA_part, (__half *)c->ws_du16, c->ws_rowscale);
  }
  TRY(launch_gemm_tc(c, "fold_gemm_dx", B, KX, 3 * c->fold_nup, c->ws_du16, c->map_g2, dx_part, KX, ksplit,
                     c->fold_inv_scale));
  return 0;
}

// Forward + backward of the folded keypoint path in one kernel, then dx = du G (see k_fold_step_w:
// all gradients are for a unit loss scale; launch_pose_bwd applies gscale / num_present).
int launch_fold_step(smplb_ctx *c, int B, const float *A, const float *cam, const float *kp_gt, float *joints,
                     float *kp_pred, float *part, int *cnt, float *d_cam, float *dA_part, float *dx_part, int ksplit) {
  RET_IF(!c->fold_ok, SMPLB_ESTATE, "folded keypoint path is not initialised");
  RET_IF(c->K * NJ * 3 > FS_MAXU, SMPLB_EINVAL, "too many keypoints for k_fold_step_w");
  LAUNCH(c, "fold_step_fwd_bwd", cdiv(B, FS_WARPS), 32 * FS_WARPS, 0, k_fold_step_w, B, c->K, c->fold_nup, c->fold_nup, c->ws_U,
         c->d_cc, A, cam, kp_gt, joints, kp_pred, part, cnt, d_cam, dA_part, (__half *)c->ws_du16, c->ws_rowscale);
  // smplb_step reduces the loss partials on stream3 from here on, next to the GEMM
  c->red_fork_recorded = false;
  if (c->ev_red_fork && c->cur == c->stream && c->use_overlap && !c->profile_serial) {
    CUDA_TRY(cudaEventRecord(c->ev_red_fork, c->cur));
    c->red_fork_recorded = true;
  }
  if (c->red_fork_recorded && c->use_prio && c->stream_g) {
    // the GEMM goes to the low-priority stream (smplb_internal.h: stream_g); smplb_step joins it before k_pose_bwd
    CUDA_TRY(cudaStreamWaitEvent(c->stream_g, c->ev_red_fork, 0));
    c->cur = c->stream_g;
    int rc = launch_gemm_tc(c, "fold_gemm_dx", B, KX, 3 * c->fold_nup, c->ws_du16, c->map_g2, dx_part, KX, ksplit,
                            c->fold_inv_scale);
    c->cur = c->stream;
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c->ev_g_join, c->stream_g));
    c->gdx_pending = true;
    return 0;
  }
  TRY(launch_gemm_tc(c, "fold_gemm_dx", B, KX, 3 * c->fold_nup, c->ws_du16, c->map_g2, dx_part, KX, ksplit,
                     c->fold_inv_scale));
  return 0;
}
