// Per-body kernels: Rodrigues, folded joint regression, the 24-joint kinematic chain and
// their backward.  One warp owns one body; lane j owns joint j.
//
// Reference semantics (paths relative to the reference checkout):
//   batch_rodrigues                    src/tf_smpl/batch_lbs.py:42-64  (eps placement :52-53)
//   batch_skew                         src/tf_smpl/batch_lbs.py:15-39
//   batch_global_rigid_transformation  src/tf_smpl/batch_lbs.py:91-152
//   joint regression + pose feature    src/tf_smpl/batch_smpl.py:115-127
#include <cuda_fp16.h>

#include "smplb_internal.h"

#define FULL 0xffffffffu

// R = cos(a) I + (1 - cos a) r r^T + sin(a) [r]x with a = ||theta + 1e-8||, r = theta / a.
// The 1e-8 is added to every component in fp32 BEFORE the norm and the numerator is the
// un-shifted theta, exactly as batch_lbs.py:52-53.  Precise sinf/cosf/div/sqrt (no fast-math).
__device__ __forceinline__ void rodrigues_fwd(float tx, float ty, float tz, float *R) {
  float ex = __fadd_rn(tx, 1e-8f), ey = __fadd_rn(ty, 1e-8f), ez = __fadd_rn(tz, 1e-8f);
  float a = sqrtf(ex * ex + ey * ey + ez * ez);
  float rx = tx / a, ry = ty / a, rz = tz / a;
  float c = cosf(a), s = sinf(a);
  float oc = 1.0f - c;
  R[0] = c + oc * rx * rx;
  R[1] = oc * rx * ry - s * rz;
  R[2] = oc * rx * rz + s * ry;
  R[3] = oc * ry * rx + s * rz;
  R[4] = c + oc * ry * ry;
  R[5] = oc * ry * rz - s * rx;
  R[6] = oc * rz * rx - s * ry;
  R[7] = oc * rz * ry + s * rx;
  R[8] = c + oc * rz * rz;
}

// d theta given G = dL/dR (validated against autograd through the
// reference by the CPU tests).
__device__ __forceinline__ void rodrigues_bwd(float tx, float ty, float tz, const float *G, float *dth) {
  float ex = __fadd_rn(tx, 1e-8f), ey = __fadd_rn(ty, 1e-8f), ez = __fadd_rn(tz, 1e-8f);
  float a = sqrtf(ex * ex + ey * ey + ez * ez);
  float ia = 1.0f / a;
  float rx = tx * ia, ry = ty * ia, rz = tz * ia;
  float c = cosf(a), s = sinf(a);
  float ax0 = G[7] - G[5], ax1 = G[2] - G[6], ax2 = G[3] - G[1];
  float Gr0 = G[0] * rx + G[1] * ry + G[2] * rz;
  float Gr1 = G[3] * rx + G[4] * ry + G[5] * rz;
  float Gr2 = G[6] * rx + G[7] * ry + G[8] * rz;
  float Tr0 = G[0] * rx + G[3] * ry + G[6] * rz;
  float Tr1 = G[1] * rx + G[4] * ry + G[7] * rz;
  float Tr2 = G[2] * rx + G[5] * ry + G[8] * rz;
  float g_c = (G[0] + G[4] + G[8]) - (rx * Gr0 + ry * Gr1 + rz * Gr2);
  float g_s = rx * ax0 + ry * ax1 + rz * ax2;
  float oc = 1.0f - c;
  float gr0 = oc * (Gr0 + Tr0) + s * ax0;
  float gr1 = oc * (Gr1 + Tr1) + s * ax1;
  float gr2 = oc * (Gr2 + Tr2) + s * ax2;
  float g_a = -s * g_c + c * g_s;
  float ux = ex * ia, uy = ey * ia, uz = ez * ia;
  float k = (tx * gr0 + ty * gr1 + tz * gr2) * ia * ia;
  float m = g_a - k;
  dth[0] = gr0 * ia + m * ux;
  dth[1] = gr1 * ia + m * uy;
  dth[2] = gr2 * ia + m * uz;
}

// Level-synchronous kinematic chain in registers.  On entry lane j (< 24) holds its local
// rotation R[9] and rest joint J[3]; on exit G[12] = (Rg row-major | tg) of its global
// transform G_j = G_parent(j) * [R_j | J_j - J_parent(j)]  (batch_lbs.py:128-135).  Parent
// transforms travel by warp shuffle; joints of depth d are composed in round d.
__device__ __forceinline__ void chain_fwd(const Tree &tree, int lane, const float *R, const float *J, float *G) {
  int j = lane < NJ ? lane : NJ - 1;
  int p = tree.parent[j];
  int psrc = p < 0 ? 0 : p;
  float pJ0 = __shfl_sync(FULL, J[0], psrc), pJ1 = __shfl_sync(FULL, J[1], psrc), pJ2 = __shfl_sync(FULL, J[2], psrc);
  float t0 = J[0], t1 = J[1], t2 = J[2];
  if (p >= 0) {
    t0 -= pJ0;
    t1 -= pJ1;
    t2 -= pJ2;
  }
#pragma unroll
  for (int e = 0; e < 9; ++e) G[e] = R[e];
  G[9] = t0;
  G[10] = t1;
  G[11] = t2;
  int myd = tree.depth[j];
  for (int d = 1; d <= tree.max_depth; ++d) {
    float P[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) P[e] = __shfl_sync(FULL, G[e], psrc);
    if (myd == d) {
      float N[12];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          N[3 * r + cc] = P[3 * r + 0] * R[cc] + P[3 * r + 1] * R[3 + cc] + P[3 * r + 2] * R[6 + cc];
        N[9 + r] = P[3 * r + 0] * t0 + P[3 * r + 1] * t1 + P[3 * r + 2] * t2 + P[9 + r];
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) G[e] = N[e];
    }
  }
}

struct __align__(16) PoseFwdSmem {
  float Rs[NJ * 9], J[NJ * 3], A[NJ * 12], Jtr[NJ * 3];
  __half x16[256], A16[12 * 64], x16b[704];
};

// bytes (a multiple of 16) from shared to global memory with 16-byte stores; dst may be null
__device__ __forceinline__ void copy_out16(int lane, void *dst, const void *src, int bytes) {
  if (!dst) return;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) != 0) {   // a caller's oddly aligned buffer: 4-byte stores
    const uint32_t *s1 = reinterpret_cast<const uint32_t *>(src);
    uint32_t *d1 = reinterpret_cast<uint32_t *>(dst);
    for (int i = lane; i < bytes / 4; i += 32) d1[i] = s1[i];
    return;
  }
  const float4 *s4 = reinterpret_cast<const float4 *>(src);
  float4 *d4 = reinterpret_cast<float4 *>(dst);
  for (int i = lane; i < bytes / 16; i += 32) d4[i] = s4[i];
}

// One warp per body: theta -> Rs, pose_feature; beta -> J (folded regression
// J = J0 + Jdirs beta, exact algebra for batch_smpl.py:110-118); chain -> A, J_transformed.
// Also writes the blend GEMM operand row x = [pose_feature | beta | 1 | 0 ...].
__global__ void __launch_bounds__(128) k_pose_fwd(int B, int NB, Tree tree, const float *__restrict__ beta,
                                                  const float *__restrict__ theta, const float *__restrict__ J0,
                                                  const float *__restrict__ Jdirs, float *__restrict__ Rs,
                                                  float *__restrict__ Jout, float *__restrict__ A,
                                                  float *__restrict__ Jtr, float *__restrict__ x,
                                                  __half *__restrict__ x16, __half *__restrict__ A16,
                                                  __half *__restrict__ x16b) {
  // Every output row of a body is staged in shared memory and written out with 16-byte stores:
  // the 2- and 4-byte stores scattered over ~6 KB per body that this kernel used to issue cost 5x
  // the DRAM write traffic of the data (partial sectors) and kept the LSU queues full (ncu).
  __shared__ PoseFwdSmem sm[4];
  int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  PoseFwdSmem &S = sm[threadIdx.x >> 5];
  int j = lane < NJ ? lane : NJ - 1;
  bool act = lane < NJ;
  const float *th = theta + (size_t)b * 72 + 3 * j;
  float R[9];
  rodrigues_fwd(th[0], th[1], th[2], R);
  float J[3];
#pragma unroll
  for (int cc = 0; cc < 3; ++cc) {
    float acc = J0[3 * j + cc];
    const float *jd = Jdirs + (size_t)(3 * j + cc) * NB;
    for (int k = 0; k < NB; ++k) acc = fmaf(jd[k], beta[(size_t)b * NB + k], acc);
    J[cc] = acc;
  }
  if (act) {
#pragma unroll
    for (int e = 0; e < 9; ++e) S.Rs[j * 9 + e] = R[e];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) S.J[j * 3 + cc] = J[cc];
    if (x && j >= 1) {
      // pose_feature index (j-1)*9 + 3r + c (batch_smpl.py:126-127)
#pragma unroll
      for (int e = 0; e < 9; ++e) x[(size_t)b * KX + (j - 1) * 9 + e] = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
    }
  }
  if (x16) {
    // fp16 operand row of the tcgen05 blend GEMM (K map in k_blend_tc.cu): pose_feature,
    // beta_hi, beta_hi, beta_lo, three 1s (v_template hi/mid/lo), zero padding to 256.
    __half *xr = S.x16;
    if (act && j >= 1) {
#pragma unroll
      for (int e = 0; e < 9; ++e)
        xr[(j - 1) * 9 + e] = __float2half_rn(R[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f));
    }
    for (int k = NPF + lane; k < 256; k += 32) {
      float v = 0.0f;
      if (k < NPF + 30) {
        int slot = (k - NPF) / 10, bi = (k - NPF) % 10;
        if (bi < NB) {
          float bv = beta[(size_t)b * NB + bi];
          float hi = __half2float(__float2half_rn(bv));
          v = (slot == 2) ? (bv - hi) : hi;
        }
      } else if (k < NPF + 33) {
        v = 1.0f;
      }
      xr[k] = __float2half_rn(v);
    }
  }
  if (x16b) {
    // operand row of the folded keypoint GEMM (k_fold.cu): x = [pose_feature | beta | 1 | 0..]
    // (224 wide) as x_hi | x_hi | x_lo | 0 (704 halves)
    __half *xr = S.x16b;
    if (act && j >= 1) {
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        float val = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
        __half hi = __float2half_rn(val);
        int k = (j - 1) * 9 + e;
        xr[k] = hi;
        xr[KX + k] = hi;
        xr[2 * KX + k] = __float2half_rn(val - __half2float(hi));
      }
    }
    for (int k = NPF + lane; k < KX; k += 32) {
      float val = 0.0f;
      if (k < NPF + NB) val = beta[(size_t)b * NB + (k - NPF)];
      else if (k == NPF + NB) val = 1.0f;
      __half hi = __float2half_rn(val);
      xr[k] = hi;
      xr[KX + k] = hi;
      xr[2 * KX + k] = __float2half_rn(val - __half2float(hi));
    }
    xr[3 * KX + lane] = __float2half_rn(0.f);
  }
  if (x) {
    // tail of the operand row: beta, the constant 1 that multiplies v_template, zero padding
    for (int k = NPF + lane; k < KX; k += 32) {
      float v = 0.0f;
      if (k < NPF + NB) v = beta[(size_t)b * NB + (k - NPF)];
      else if (k == NPF + NB) v = 1.0f;
      x[(size_t)b * KX + k] = v;
    }
  }
  float G[12];
  chain_fwd(tree, lane, R, J, G);
  if (act) {
    // A_j = [Rg | tg - Rg J_j]  (batch_lbs.py:146-150), J_transformed = tg (:140)
    float *a = S.A + j * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      a[4 * r + 0] = G[3 * r + 0];
      a[4 * r + 1] = G[3 * r + 1];
      a[4 * r + 2] = G[3 * r + 2];
      a[4 * r + 3] = G[9 + r] - (G[3 * r + 0] * J[0] + G[3 * r + 1] * J[1] + G[3 * r + 2] * J[2]);
    }
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) S.Jtr[j * 3 + cc] = G[9 + cc];
    if (A16) {
      // fp16 split operand of the tcgen05 skinning GEMM (k_skin_tc.cu): row (b, e = 4r + d), one
      // 64-column swizzle atom laid out in 16-column windows (see the window table there):
      //   [A_hi 0..15] [A_hi 16..23 | A_lo 16..23] [A_lo 0..15] [0]
      __half *rowbase = S.A16;
      const int col_lo = j < 16 ? 32 + j : j + 8;
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        float val = a[e];
        __half hi = __float2half_rn(val);
        __half lo = __float2half_rn(val - __half2float(hi));
        rowbase[e * 64 + j] = hi;
        rowbase[e * 64 + col_lo] = lo;
      }
    }
  }
  if (A16) {
    // zero padding columns 48..63 of the 12 operand rows
    for (int i = lane; i < 12 * 16; i += 32) S.A16[(i / 16) * 64 + 48 + (i % 16)] = __float2half_rn(0.f);
  }
  __syncwarp();
  copy_out16(lane, Rs ? Rs + (size_t)b * NJ * 9 : nullptr, S.Rs, NJ * 9 * 4);
  copy_out16(lane, Jout ? Jout + (size_t)b * NJ * 3 : nullptr, S.J, NJ * 3 * 4);
  copy_out16(lane, A + (size_t)b * NJ * 12, S.A, NJ * 12 * 4);
  copy_out16(lane, Jtr ? Jtr + (size_t)b * NJ * 3 : nullptr, S.Jtr, NJ * 3 * 4);
  copy_out16(lane, x16 ? x16 + (size_t)b * 256 : nullptr, S.x16, 256 * 2);
  copy_out16(lane, A16 ? A16 + (size_t)b * 12 * 64 : nullptr, S.A16, 12 * 64 * 2);
  copy_out16(lane, x16b ? x16b + (size_t)b * 704 : nullptr, S.x16b, 704 * 2);
}

// Backward of the per-body stage.  One warp per body; the reverse chain walks joints 23..1
// with the per-joint state in shared memory (children always have larger indices than their
// parent, batch_lbs.py:128).  Inputs: fixed-order partial sums of dL/dA from the skinning
// backward and of dL/dx from the blend backward GEMM.
#define PB_WARPS 4
struct PoseBwdSmem {
  float Rg[NJ][9], R[NJ][9], J[NJ][3];
  float dRg[NJ][9], dR[NJ][9], dtg[NJ][3], dJ[NJ][3];
};

__global__ void __launch_bounds__(32 * PB_WARPS)
    k_pose_bwd(int B, int NB, Tree tree, const float *__restrict__ theta, const float *__restrict__ Rs,
               const float *__restrict__ Jin, const float *__restrict__ A, const float *__restrict__ dA_part,
               int n_dA_parts, const float *__restrict__ dx_part, int ksplit, int dx_rows,
               const float *__restrict__ rowscale, const float *__restrict__ d_Rs,
               const float *__restrict__ Jdirs, float *__restrict__ d_beta, float *__restrict__ d_theta,
               const long long *__restrict__ den, float gscale, float *__restrict__ d_cam) {
  __shared__ PoseBwdSmem sm[PB_WARPS];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x * PB_WARPS + w;
  if (b >= B) return;
  PoseBwdSmem &S = sm[w];
  // Deferred loss scale (k_fold_step_w): dA, dx and d_cam arrive for a unit upstream scale and
  // everything below is linear in them, so gscale / num_present is applied to the outputs.
  float osc = 1.0f;
  if (den) {
    long long dv = *den;
    osc = dv > 0 ? gscale / (float)dv : 0.0f;
    if (d_cam && lane < 3) d_cam[(size_t)b * 3 + lane] *= osc;
  }
  if (lane < NJ) {
    int j = lane;
    float dA[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) dA[e] = 0.0f;
    for (int sp = 0; sp < n_dA_parts; ++sp) {
      const float *src = dA_part + ((size_t)sp * B + b) * (NJ * 12) + j * 12;
#pragma unroll
      for (int e = 0; e < 12; ++e) dA[e] += src[e];
    }
    const float *a = A + ((size_t)b * NJ + j) * 12;
    float J[3];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      J[cc] = Jin[((size_t)b * NJ + j) * 3 + cc];
      S.J[j][cc] = J[cc];
    }
    float Rg[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        Rg[3 * r + cc] = a[4 * r + cc];
        S.Rg[j][3 * r + cc] = Rg[3 * r + cc];
        S.R[j][3 * r + cc] = Rs[((size_t)b * NJ + j) * 9 + 3 * r + cc];
        // dRg = dA_R - dA_t (x) J
        S.dRg[j][3 * r + cc] = dA[4 * r + cc] - dA[4 * r + 3] * J[cc];
        S.dR[j][3 * r + cc] = 0.0f;
      }
#pragma unroll
    for (int r = 0; r < 3; ++r) S.dtg[j][r] = dA[4 * r + 3];
    // dJ = -Rg^T dA_t
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
      S.dJ[j][cc] = -(Rg[cc] * dA[3] + Rg[3 + cc] * dA[7] + Rg[6 + cc] * dA[11]);
  }
  __syncwarp();
  for (int i = NJ - 1; i >= 1; --i) {
    int p = tree.parent[i];
    if (p < 0) continue;  // a second root: nothing to propagate
    if (lane < 9) {
      int r = lane / 3, cc = lane % 3;
      // dR_i = Rg_p^T dRg_i
      S.dR[i][lane] = S.Rg[p][0 + r] * S.dRg[i][0 + cc] + S.Rg[p][3 + r] * S.dRg[i][3 + cc] + S.Rg[p][6 + r] * S.dRg[i][6 + cc];
      // dRg_p += dRg_i R_i^T + dtg_i (x) (J_i - J_p)
      float add = S.dRg[i][3 * r + 0] * S.R[i][3 * cc + 0] + S.dRg[i][3 * r + 1] * S.R[i][3 * cc + 1] +
                  S.dRg[i][3 * r + 2] * S.R[i][3 * cc + 2] + S.dtg[i][r] * (S.J[i][cc] - S.J[p][cc]);
      S.dRg[p][lane] += add;
    } else if (lane < 12) {
      int r = lane - 9;
      float djr = S.Rg[p][0 + r] * S.dtg[i][0] + S.Rg[p][3 + r] * S.dtg[i][1] + S.Rg[p][6 + r] * S.dtg[i][2];
      S.dJ[i][r] += djr;
      S.dJ[p][r] -= djr;
      S.dtg[p][r] += S.dtg[i][r];
    }
    __syncwarp();
  }
  // roots: dR = dRg, dJ += dtg
  if (lane < NJ && tree.parent[lane] < 0) {
#pragma unroll
    for (int e = 0; e < 9; ++e) S.dR[lane][e] = S.dRg[lane][e];
#pragma unroll
    for (int r = 0; r < 3; ++r) S.dJ[lane][r] += S.dtg[lane][r];
  }
  __syncwarp();
  if (lane < NJ) {
    int j = lane;
    float G[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      float g = S.dR[j][e];
      if (j >= 1) {
        float acc = 0.0f;
        for (int ks = 0; ks < ksplit; ++ks) acc += dx_part[((size_t)ks * dx_rows + b) * KX + (j - 1) * 9 + e];
        g += rowscale ? acc * rowscale[b] : acc;
      }
      if (d_Rs) g += d_Rs[((size_t)b * NJ + j) * 9 + e];
      G[e] = g;
    }
    const float *th = theta + (size_t)b * 72 + 3 * j;
    float dth[3];
    rodrigues_bwd(th[0], th[1], th[2], G, dth);
    d_theta[(size_t)b * 72 + 3 * j + 0] = dth[0] * osc;
    d_theta[(size_t)b * 72 + 3 * j + 1] = dth[1] * osc;
    d_theta[(size_t)b * 72 + 3 * j + 2] = dth[2] * osc;
  }
  if (lane < NB) {
    // d beta = Jdirs^T dJ + (dp . shapedirs^T), the latter from the blend backward GEMM
    float acc = 0.0f;
    for (int ks = 0; ks < ksplit; ++ks) acc += dx_part[((size_t)ks * dx_rows + b) * KX + NPF + lane];
    if (rowscale) acc *= rowscale[b];
    for (int jc = 0; jc < NJ * 3; ++jc) acc = fmaf(Jdirs[(size_t)jc * NB + lane], S.dJ[jc / 3][jc % 3], acc);
    d_beta[(size_t)b * NB + lane] = acc * osc;
  }
}

// The same backward without the 23-step serial walk: everything the reverse chain accumulates is a SUBTREE SUM.
// With S_i = sum over the subtree of i of dtg0 (dtg0 = dA_t), Rg_k = Rg_i * (product of the local rotations from i
// down to k), the recursion  dRg_p += dRg_i R_i^T + dtg_i (x) (J_i - J_p),  dtg_p += dtg_i  unrolls to
//     dRg_i = (sum over k in subtree(i) of E_k Rg_k^T) Rg_i,   E_k = dRg0_k + sum over children c of k of S_c (x) (J_c - J_k),
// and   dR_i = Rg_p^T dRg_i,   dJ_i = dJ0_i + Rg_p^T S_i - Rg_i^T (S_i - dtg0_i)   (roots: dR = dRg, Rg_p = I).
// One warp per body, lane = joint, state in registers; subtree sums gather children level by level with shuffles
// (children in index order, so the sums are deterministic).  No shared memory, no local rotations.
// ch[slot] = this lane's slot-th child (-1: none), slots[d] = the largest child count among the joints of depth d - 1
template <int N>
__device__ __forceinline__ void subtree_sum(float *v, const int *ch, int depth, int max_depth, unsigned slots) {
#pragma unroll 1
  for (int d = max_depth; d >= 1; --d) {
    const int ns = (slots >> (2 * d)) & 3;
#pragma unroll 1
    for (int slot = 0; slot < ns; ++slot) {
      const int cs = slot == 0 ? ch[0] : (slot == 1 ? ch[1] : ch[2]);
      const bool on = depth == d - 1 && cs >= 0;
      const int c = on ? cs : 0;
#pragma unroll
      for (int e = 0; e < N; ++e) {
        const float x = __shfl_sync(FULL, v[e], c);
        v[e] += on ? x : 0.0f;
      }
    }
  }
}

__global__ void __launch_bounds__(128)
    k_pose_bwd_reg(int B, int NB, Tree tree, const float *__restrict__ theta, const float *__restrict__ Jin,
                   const float *__restrict__ A, const float *__restrict__ dA_part, int n_dA_parts,
                   const float *__restrict__ dx_part, int ksplit, int dx_rows, const float *__restrict__ rowscale,
                   const float *__restrict__ d_Rs, const float *__restrict__ Jdirs, float *__restrict__ d_beta,
                   float *__restrict__ d_theta, const long long *__restrict__ den, float gscale, float *__restrict__ d_cam) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float osc = 1.0f;
  if (den) {
    long long dv = *den;
    osc = dv > 0 ? gscale / (float)dv : 0.0f;
    if (d_cam && lane < 3) d_cam[(size_t)b * 3 + lane] *= osc;
  }
  const bool act = lane < NJ;
  const int j = act ? lane : NJ - 1;
  const int par = act ? tree.parent[j] : -2;
  const int depth = act ? tree.depth[j] : -1;
  const int ch[3] = {act ? tree.child[j][0] : -1, act ? tree.child[j][1] : -1, act ? tree.child[j][2] : -1};
  const unsigned slots = tree.level_slots;
  // independent loads first: dx partials (pose-feature rows), theta, then dA / A / J
  float gx[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    float acc = 0.0f;
    if (j >= 1)
      for (int ks = 0; ks < ksplit; ++ks) acc += dx_part[((size_t)ks * dx_rows + b) * KX + (j - 1) * 9 + e];
    gx[e] = acc;
  }
  const float rs = rowscale ? rowscale[b] : 1.0f;
  float dA[12];
  {
    const float4 *src = reinterpret_cast<const float4 *>(dA_part + (size_t)b * (NJ * 12) + j * 12);
    float4 a0 = src[0], a1 = src[1], a2 = src[2];
    for (int sp = 1; sp < n_dA_parts; ++sp) {
      const float4 *s2 = src + (size_t)sp * B * (NJ * 3);
      float4 b0 = s2[0], b1 = s2[1], b2 = s2[2];
      a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
      a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
      a2.x += b2.x; a2.y += b2.y; a2.z += b2.z; a2.w += b2.w;
    }
    dA[0] = a0.x; dA[1] = a0.y; dA[2] = a0.z; dA[3] = a0.w;
    dA[4] = a1.x; dA[5] = a1.y; dA[6] = a1.z; dA[7] = a1.w;
    dA[8] = a2.x; dA[9] = a2.y; dA[10] = a2.z; dA[11] = a2.w;
  }
  float Rg[9], J[3];
  {
    const float4 *src = reinterpret_cast<const float4 *>(A + ((size_t)b * NJ + j) * 12);
    float4 a0 = src[0], a1 = src[1], a2 = src[2];
    Rg[0] = a0.x; Rg[1] = a0.y; Rg[2] = a0.z;
    Rg[3] = a1.x; Rg[4] = a1.y; Rg[5] = a1.z;
    Rg[6] = a2.x; Rg[7] = a2.y; Rg[8] = a2.z;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) J[cc] = Jin[((size_t)b * NJ + j) * 3 + cc];
  }
  if (!act) {
#pragma unroll
    for (int e = 0; e < 12; ++e) dA[e] = 0.0f;
  }
  const float t0[3] = {dA[3], dA[7], dA[11]};            // dtg0 = dA_t
  float S[3] = {t0[0], t0[1], t0[2]};
  subtree_sum<3>(S, ch, depth, tree.max_depth, slots);
  // Z = sum over children c of S_c (x) J_c
  float Z[9];
  {
    float Y[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        Y[3 * r + cc] = S[r] * J[cc];
        Z[3 * r + cc] = 0.0f;
      }
#pragma unroll
    for (int slot = 0; slot < 3; ++slot) {
      const bool on = ch[slot] >= 0;
      const int c = on ? ch[slot] : 0;
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        const float x = __shfl_sync(FULL, Y[e], c);
        Z[e] += on ? x : 0.0f;
      }
    }
  }
  // E = dRg0 + Z - (S - dtg0) (x) J with dRg0 = dA_R - dA_t (x) J;  Q = subtree sum of E Rg^T
  float Q[9];
  {
    float E[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) E[3 * r + cc] = (dA[4 * r + cc] - t0[r] * J[cc]) + (Z[3 * r + cc] - (S[r] - t0[r]) * J[cc]);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) Q[3 * r + cc] = E[3 * r] * Rg[3 * cc] + E[3 * r + 1] * Rg[3 * cc + 1] + E[3 * r + 2] * Rg[3 * cc + 2];
  }
  subtree_sum<9>(Q, ch, depth, tree.max_depth, slots);
  // dRg = Q Rg; parent's rotation by shuffle (identity for a root)
  float dRg[9], P[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) dRg[3 * r + cc] = Q[3 * r] * Rg[cc] + Q[3 * r + 1] * Rg[3 + cc] + Q[3 * r + 2] * Rg[6 + cc];
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    const float x = __shfl_sync(FULL, Rg[e], par >= 0 ? par : 0);
    P[e] = par >= 0 ? x : ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
  }
  // G = dR (+ blend-backward pose-feature gradient + upstream d_Rs);  dR = Rg_p^T dRg
  float G[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      float g = P[r] * dRg[cc] + P[3 + r] * dRg[3 + cc] + P[6 + r] * dRg[6 + cc];
      g += gx[3 * r + cc] * rs;
      if (d_Rs) g += d_Rs[((size_t)b * NJ + j) * 9 + 3 * r + cc];
      G[3 * r + cc] = g;
    }
  // dJ = dJ0 + Rg_p^T S - Rg^T (S - dtg0),  dJ0 = -Rg^T dtg0   =>   dJ = Rg_p^T S - Rg^T S
  float dJ[3];
#pragma unroll
  for (int cc = 0; cc < 3; ++cc)
    dJ[cc] = (P[cc] * S[0] + P[3 + cc] * S[1] + P[6 + cc] * S[2]) - (Rg[cc] * S[0] + Rg[3 + cc] * S[1] + Rg[6 + cc] * S[2]);
  if (act) {
    const float *th = theta + (size_t)b * 72 + 3 * j;
    float dth[3];
    rodrigues_bwd(th[0], th[1], th[2], G, dth);
    d_theta[(size_t)b * 72 + 3 * j + 0] = dth[0] * osc;
    d_theta[(size_t)b * 72 + 3 * j + 1] = dth[1] * osc;
    d_theta[(size_t)b * 72 + 3 * j + 2] = dth[2] * osc;
  } else {
    dJ[0] = dJ[1] = dJ[2] = 0.0f;
  }
  // d beta = Jdirs^T dJ (+ dp . shapedirs^T from the blend backward GEMM): lane j adds its joint's three rows of
  // Jdirs, then a fixed-order butterfly over the lanes
  float db[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) db[k] = 0.0f;
  if (act) {
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      const float *jd = Jdirs + (size_t)(3 * j + cc) * NB;
#pragma unroll
      for (int k = 0; k < 10; ++k)
        if (k < NB) db[k] = fmaf(jd[k], dJ[cc], db[k]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 10; ++k) db[k] += __shfl_xor_sync(FULL, db[k], o);
  if (lane < NB) {
    float acc = 0.0f;
    for (int ks = 0; ks < ksplit; ++ks) acc += dx_part[((size_t)ks * dx_rows + b) * KX + NPF + lane];
    acc *= rs;
    float mine = db[0];
#pragma unroll
    for (int k = 1; k < 10; ++k)
      if (lane == k) mine = db[k];
    d_beta[(size_t)b * NB + lane] = (acc + mine) * osc;
  }
}

// batch_rodrigues stand-alone (batch_lbs.py:42): one thread per rotation.
__global__ void k_rodrigues(int N, const float *__restrict__ theta, float *__restrict__ R) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float r[9];
  rodrigues_fwd(theta[3 * (size_t)i], theta[3 * (size_t)i + 1], theta[3 * (size_t)i + 2], r);
#pragma unroll
  for (int e = 0; e < 9; ++e) R[(size_t)i * 9 + e] = r[e];
}

// batch_global_rigid_transformation stand-alone (batch_lbs.py:91): caller-supplied Rs, Js;
// A is the reference's [B,24,4,4] with bottom rows [0,0,0,1].
__global__ void __launch_bounds__(128) k_global_rigid(int B, Tree tree, const float *__restrict__ Rs,
                                                      const float *__restrict__ Js, float *__restrict__ newJ,
                                                      float *__restrict__ A44) {
  int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  int j = lane < NJ ? lane : NJ - 1;
  float R[9], J[3], G[12];
#pragma unroll
  for (int e = 0; e < 9; ++e) R[e] = Rs[((size_t)b * NJ + j) * 9 + e];
#pragma unroll
  for (int cc = 0; cc < 3; ++cc) J[cc] = Js[((size_t)b * NJ + j) * 3 + cc];
  chain_fwd(tree, lane, R, J, G);
  if (lane < NJ) {
    float *a = A44 + ((size_t)b * NJ + j) * 16;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      a[4 * r + 0] = G[3 * r + 0];
      a[4 * r + 1] = G[3 * r + 1];
      a[4 * r + 2] = G[3 * r + 2];
      a[4 * r + 3] = G[9 + r] - (G[3 * r + 0] * J[0] + G[3 * r + 1] * J[1] + G[3 * r + 2] * J[2]);
      newJ[((size_t)b * NJ + j) * 3 + r] = G[9 + r];
    }
    a[12] = 0.0f;
    a[13] = 0.0f;
    a[14] = 0.0f;
    a[15] = 1.0f;
  }
}

// batch_skew stand-alone (batch_lbs.py:15-39): [[0,-z,y],[z,0,-x],[-y,x,0]].
__global__ void k_skew(int N, const float *__restrict__ vec, float *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float x = vec[3 * (size_t)i], y = vec[3 * (size_t)i + 1], z = vec[3 * (size_t)i + 2];
  float *o = out + (size_t)i * 9;
  o[0] = 0.f; o[1] = -z; o[2] = y;
  o[3] = z; o[4] = 0.f; o[5] = -x;
  o[6] = -y; o[7] = x; o[8] = 0.f;
}

// batch_lrotmin stand-alone (batch_lbs.py:67-88): (Rodrigues(theta[:, 3:]) - I) -> [B,207].
__global__ void k_lrotmin(int B, const float *__restrict__ theta, float *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 23) return;
  int b = i / 23, j = i % 23 + 1;
  float r[9];
  const float *th = theta + (size_t)b * 72 + 3 * j;
  rodrigues_fwd(th[0], th[1], th[2], r);
#pragma unroll
  for (int e = 0; e < 9; ++e) out[(size_t)b * NPF + (j - 1) * 9 + e] = r[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
}

int launch_skew(smplb_ctx *c, int N, const float *vec, float *out) {
  LAUNCH(c, "skew", cdiv(N, 128), 128, 0, k_skew, N, vec, out);
  return 0;
}

int launch_lrotmin(smplb_ctx *c, int B, const float *theta, float *out) {
  LAUNCH(c, "lrotmin", cdiv(B * 23, 128), 128, 0, k_lrotmin, B, theta, out);
  return 0;
}

int launch_pose_fwd(smplb_ctx *c, int B, const float *beta, const float *theta, float *Rs, float *J, float *A,
                    float *Jtr, float *x, void *x16, void *A16, void *x16b) {
  LAUNCH(c, "pose_fwd", cdiv(B, 4), 128, 0, k_pose_fwd, B, c->NB, c->tree, beta, theta, c->d_J0, c->d_Jdirs, Rs, J, A,
         Jtr, x, (__half *)x16, (__half *)A16, (__half *)x16b);
  return 0;
}

int launch_pose_bwd(smplb_ctx *c, int B, const float *theta, const float *Rs, const float *J, const float *A,
                    const float *dA_part, int n_dA_parts, const float *dx_part, int ksplit, int dx_rows,
                    const float *rowscale, const float *d_Rs, float *d_beta, float *d_theta, const long long *den,
                    float gscale, float *d_cam) {
  if (c->use_pose_bwd_reg && c->NB <= 10 && c->tree.max_children <= 3 && c->tree.max_depth < 16) {
    // register / shuffle formulation (subtree sums); smplb_debug_set("pose_bwd_reg", 0) selects the serial walk
    LAUNCH(c, "pose_bwd", cdiv(B, 4), 128, 0, k_pose_bwd_reg, B, c->NB, c->tree, theta, J, A, dA_part, n_dA_parts, dx_part, ksplit,
           dx_rows, rowscale, d_Rs, c->d_Jdirs, d_beta, d_theta, den, gscale, d_cam);
    return 0;
  }
  LAUNCH(c, "pose_bwd", cdiv(B, PB_WARPS), 32 * PB_WARPS, 0, k_pose_bwd, B, c->NB, c->tree, theta, Rs, J, A, dA_part,
         n_dA_parts, dx_part, ksplit, dx_rows, rowscale, d_Rs, c->d_Jdirs, d_beta, d_theta, den, gscale, d_cam);
  return 0;
}

int launch_rodrigues(smplb_ctx *c, int N, const float *theta, float *R) {
  LAUNCH(c, "rodrigues", cdiv(N, 128), 128, 0, k_rodrigues, N, theta, R);
  return 0;
}

int launch_global_rigid(smplb_ctx *c, int B, const float *Rs, const float *Js, float *new_J, float *A44) {
  LAUNCH(c, "global_rigid", cdiv(B, 4), 128, 0, k_global_rigid, B, c->tree, Rs, Js, new_J, A44);
  return 0;
}
