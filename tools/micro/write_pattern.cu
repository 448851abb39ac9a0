// Microbenchmark: write-only bandwidth for blend_tc's store pattern (tiles of R rows x 512 B
// at a given row pitch) versus a linear memset, to see whether the row pitch matters.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_tiles(float4 *out, size_t pitch4, int rows, int ncolblk, int rows_per_tile) {
  // persistent: CTA walks tiles (m, n) n-fastest; tile = rows_per_tile rows x 32 float4 (512 B)
  int n_mblk = rows / rows_per_tile;
  int total = n_mblk * ncolblk;
  int t0 = (int)((long long)blockIdx.x * total / gridDim.x), t1 = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
  float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int t = t0; t < t1; ++t) {
    int m = t / ncolblk, n = t % ncolblk;
    for (int i = threadIdx.x; i < rows_per_tile * 32; i += blockDim.x) {
      int r = i / 32, c4 = i % 32;
      __stcs(out + (size_t)(m * rows_per_tile + r) * pitch4 + n * 32 + c4, v);
    }
  }
}
int main() {
  const int rows = 4096;
  for (int pad : {0, 32, 96, 544}) {
    size_t pitch = 20736 + pad;           // floats
    float4 *buf;
    cudaMalloc(&buf, pitch * 4 * rows);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rpt : {128, 256}) {
      for (int it = 0; it < 3; ++it) k_tiles<<<148, 256>>>(buf, pitch / 4, rows, 162, rpt);
      cudaEventRecord(e0);
      for (int it = 0; it < 10; ++it) k_tiles<<<148, 256>>>(buf, pitch / 4, rows, 162, rpt);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
      printf("pad %4d floats, %3d-row tiles: %.1f us -> %.0f GB/s\n", pad, rpt, ms * 1e3, 4096.0 * 20736 * 4 / ms / 1e6);
    }
    cudaFree(buf);
  }
  return 0;
}
