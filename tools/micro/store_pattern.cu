// Microbenchmark: how fast can ONE CTA per SM with 8 (or 16) warps write verts [B][V][3] fp32 when a thread
// owns one vertex (x, y, z in registers) -- the situation of the fused kernel's epilogue.
//   A: three 4-byte stores per vertex (12 B stride across the warp)            [what the kernel does]
//   B: lane pairs: one shuffle, then STG.64 on all lanes + STG.64 on the even lanes (rows are 8-byte aligned)
//   C: 4 B stores but a warp's three stores of a row each cover 128 contiguous bytes (WRONG order; coalescing bound)
//   F: 16-byte stores, fully coalesced, ignoring the layout (upper bound of the store path)
// Build: nvcc -arch=sm_100a -O3 --cudart=shared -o store_pattern store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
#define V 6890
#define NS 96
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *out, int B, int n_vt, int n_m, int order, int l2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int total = n_vt * n_m;
  const int t0 = (int)((long long)blockIdx.x * total / gridDim.x), t1 = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
  const int q = warp & 3, part = warp >> 2, nparts = nw >> 2;   // quarter of the 128-vertex tile; sample split
  // order 2: vertex-major ranges walked in a ROTATED order so that all CTAs sweep the sample blocks in step:
  // start at the first tile of the range whose sample block is a multiple of n_m / 3
  int rot = 0;
  const int n = t1 - t0;
  if (order == 2) {
    const float step = n_m / 3.0f;
    for (int i = 0; i < n; ++i) {
      const int m = (t0 + i) % n_m;
      const float r = fmodf((float)m, step);
      if (r < 1.0f) { rot = i; break; }
    }
  }
  for (int i = 0; i < n; ++i) {
    int t = t0 + (i + rot) % n;
    const int vt = order == 1 ? t % n_vt : t / n_m, m = order == 1 ? t / n_vt : t % n_m;   // order 1: sample-block major
    const int v = vt * 128 + 32 * q + lane;
    const bool ok = v < V;
    for (int s = part; s < NS; s += nparts) {
      const int b = m * NS + s;
      if (b >= B) break;
      float x = __int_as_float(b * 3 + v), y = x + 1.f, z = x + 2.f;
      float *row = out + (size_t)(l2 ? s : b) * V * 3;
      if (MODE == 0) {
        if (ok) { __stcs(row + 3 * v, x); __stcs(row + 3 * v + 1, y); __stcs(row + 3 * v + 2, z); }
      } else if (MODE == 1) {
        const float send = (lane & 1) ? x : z;
        const float got = __shfl_xor_sync(0xffffffffu, send, 1);   // even: x of the odd neighbour; odd: z of the even one
        if (ok) {
          float2 a = (lane & 1) ? make_float2(y, z) : make_float2(x, y);
          float *pa = (lane & 1) ? row + 3 * v + 1 : row + 3 * v;
          __stcs((float2 *)pa, a);
          if (!(lane & 1) && v + 1 < V) __stcs((float2 *)(row + 3 * v + 2), make_float2(z, got));
          else if (!(lane & 1)) __stcs(row + 3 * v + 2, z);
        }
      } else if (MODE == 2) {
        float *p = row + 3 * (v - lane) + lane;
        if (ok) { __stcs(p, x); __stcs(p + 32, y); __stcs(p + 64, z); }
      } else {
        // 128 vertices x 3 floats = 96 float4 per (tile, sample): warps of a quarter... just stream float4
        float4 *p4 = (float4 *)(out + ((size_t)(l2 ? s : b) * V * 3 / 4) * 4) + (vt * 96 + q * 24);
        if (lane < 24 && vt * 128 + 128 <= V) __stcs(p4 + lane, make_float4(x, y, z, x));
      }
    }
  }
}
template <int MODE>
void run(const char *name, float *buf, int B, int threads, int order, int l2 = 0) {
  int n_vt = (V + 127) / 128, n_m = (B + NS - 1) / NS;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) k<MODE><<<148, threads>>>(buf, B, n_vt, n_m, order, l2);
  cudaEventRecord(e0);
  for (int it = 0; it < 10; ++it) k<MODE><<<148, threads>>>(buf, B, n_vt, n_m, order, l2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
  printf("%s%s %-28s %2d warps: %.1f us -> %.0f GB/s  (%s)\n", l2 ? "[L2-resident] " : "", order == 1 ? "sample-major" : order == 2 ? "vertex-major rotated" : "vertex-major", name, threads / 32, ms * 1e3, (double)B * V * 12 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int B = 4096;
  float *buf;
  cudaMalloc(&buf, (size_t)B * V * 12 + 256);
  for (int order : {0, 1, 2})
    for (int th : {256, 512}) {
      run<0>("A 3 x STG.32 per vertex", buf, B, th, order);
      run<2>("C 3 x STG.32 coalesced", buf, B, th, order);
      run<3>("F STG.128 (24 lanes)", buf, B, th, order);
    }
  for (int th : {256, 512}) {
    run<0>("A 3 x STG.32 per vertex", buf, B, th, 0, 1);
    run<1>("B lane pairs, STG.64", buf, B, th, 0, 1);
    run<2>("C 3 x STG.32 coalesced", buf, B, th, 0, 1);
    run<3>("F STG.128 (24 lanes)", buf, B, th, 0, 1);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int it = 0; it < 10; ++it) cudaMemsetAsync(buf, 1, (size_t)B * V * 12);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("memset: %.1f us\n", ms * 100);
  return 0;
}
