// Issue rate of cta_group::2 MMAs (256 x N x 16, fp16) on a CTA pair: one or two issuing warps in
// the leader, `nacc` alternating accumulators.  Operand contents are irrelevant (only timing).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pair_rate pair_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../../human-pose-estimation_b200/csrc/tc_ptx.cuh"

template <int N>
__global__ void __launch_bounds__(128, 1) k_pair(int iters, int nissue, int kstep, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bar[4];
  const int warp = threadIdx.x >> 5;
  const bool leader = cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tbase = tptr;
  long long t0 = clock64();
  if (leader && warp < nissue) {
    const uint32_t sb = smem_u32(smem);
    constexpr uint32_t idesc = umma_idesc_f16(256, N);
    const uint64_t bd0 = umma_desc_sw128(sb + 16384);
    const uint64_t ad0 = umma_desc_sw128(sb);
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t d = tbase + warp * N;
        const int k = kstep ? (u & 3) : 0;
        if (elect_one()) tc_mma_f16_pair(d, ad0 + 2 * k, bd0 + 2 * k, idesc, 1);
        __syncwarp();
      }
    }
    if (elect_one()) tc_commit_pair(smem_u32(&bar[warp]));
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
  }
}

template <int N>
static void run(long long *d_out, int nissue, int kstep) {
  long long h[148];
  const int iters = 4000;
  cudaFuncSetAttribute(k_pair<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 64 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchKernelEx(&cfg, k_pair<N>, iters, nissue, kstep, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("pair mma N=%d: %s\n", N, cudaGetErrorString(e));
      return;
    }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; i += 2) mx = h[i] > mx ? h[i] : mx;
  printf("pair mma 256x%dx16, %d issuing warp(s), k-step %d: %.1f clk per MMA per warp (dense-rate floor %d)\n", N, nissue,
         kstep, (double)mx / iters, N / 2);
}

int main() {
  long long *d_out;
  cudaMalloc(&d_out, 2048 * sizeof(long long));
  run<96>(d_out, 1, 1);
  run<96>(d_out, 1, 0);
  run<96>(d_out, 2, 1);
  run<128>(d_out, 1, 1);
  run<192>(d_out, 1, 1);
  run<256>(d_out, 1, 1);
  run<64>(d_out, 1, 1);
  run<32>(d_out, 1, 1);
  return 0;
}
