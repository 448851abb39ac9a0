// Shared by the loss kernels (k_loss.cu) and the lattice nearest-neighbour search (k_mesh_lattice.cu).
#pragma once
#include "smplb_internal.h"

// Fixed-order block sum of one float per thread (blockDim.x a power of two <= 1024).
__device__ __forceinline__ float block_sum(float v, float *red) {
  int t = threadIdx.x;
  red[t] = v;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  float r = red[0];
  __syncthreads();
  return r;
}

// d2(a, b) exactly as the fp32 expansion of ops.py:63-65: (-2 a.b + |a|^2) + |b|^2, each sum
// rounded separately (-2 * x is exact, so the FMA below rounds once like the TF add does).
__device__ __forceinline__ float d2_expand(float ax, float ay, float a2, float bx, float by, float b2) {
  float dot = __fmaf_rn(ax, bx, __fmul_rn(ay, by));
  return __fadd_rn(__fmaf_rn(-2.0f, dot, a2), b2);
}

#define GP_STRIDE 8   // per (image, set) grid parameters: x0, y0, x1, y1, inv_h, h, max |p|^2, -

#define LAT_N 256                       // lattice extent (the reference's images are 224 x 224)
#define LAT_W (LAT_N / 32)              // 32-bit words per row
#define LAT_C 64                        // coarse cells per side (4 x 4 pixels each)
#define LAT_BM_WORDS (LAT_N * LAT_W)
#define LAT_BYTES (2 * LAT_BM_WORDS * 4 + LAT_N * 4 + LAT_C * LAT_C)   // bitmap, transposed bitmap, row prefix, coarse distances


