// Micro-benchmarks behind the fused blend+skin design:
//   1. tcgen05.ld throughput per SM (how fast the epilogue can drain TMEM),
//   2. tcgen05.mma issue rate at small N with the A operand in TMEM (TS mode) and in shared
//      memory (SS mode).
// Operand contents are whatever the memories hold: only the timing is read.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rate tmem_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../../human-pose-estimation_b200/csrc/tc_ptx.cuh"

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode 0: tcgen05.ld x32 from `nwarps` warps; mode 1: x16
__global__ void __launch_bounds__(256, 1) k_ld(int iters, int mode, long long *out) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tptr + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[32];
    if (mode == 0) {
      tc_ld_32x32(tbase + ((i * 32) & 255) + (warp >> 2) * 256, r);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    } else {
      tc_ld_32x16(tbase + ((i * 16) & 255) + (warp >> 2) * 256, r);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r[j];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) out[1000] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "n"(512) : "memory");
  }
}

// iters MMAs of shape 128 x N x 16, A from TMEM (ts = 1) or shared memory (ts = 0)
template <int N>
__global__ void __launch_bounds__(128, 1) k_mma(int iters, int ts, int nacc, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tptr;
  if (warp == 0) {
    // the whole warp runs the loop (warp-uniform operands stay in uniform registers); one elected
    // lane issues, as CUTLASS does
    const uint32_t sb = smem_u32(smem);
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t bd0 = umma_desc_sw128(sb + 16384);
    const uint64_t ad0 = umma_desc_sw128(sb);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t d = tbase + 128 + (nacc == 1 ? 0 : (u % 2) * N) + (nacc > 2 ? (u / 2 % 2) * 2 * N : 0);
        uint64_t bdesc = bd0 + 2 * (u & 3);
        if (elect_one()) {
          if (ts)
            tc_mma_f16_ts(d, tbase + u * 8, bdesc, idesc, 1);
          else
            tc_mma_f16(d, ad0 + 2 * (u & 3), bdesc, idesc, 1);
        }
        __syncwarp();
      }
    }
    if (elect_one()) tc_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
  }
}

template <int N>
static void run_mma(long long *d_out, int ts, int nacc = 2) {
  long long h[148];
  const int iters = 4000;
  cudaFuncSetAttribute(k_mma<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k_mma<N><<<148, 128, 64 * 1024>>>(iters, ts, nacc, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mma N=%d ts=%d: %s\n", N, ts, cudaGetErrorString(e));
      return;
    }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("mma 128x%dx16 %s nacc=%d: %.1f clk/instr (floor %d)\n", N, ts ? "TS" : "SS", nacc, (double)mx / iters, N / 2);
}

// Several warps issuing independent MMA streams: is the ~100 clk/instruction floor a property of
// the issuing thread or of the tensor pipe?
template <int N>
__global__ void __launch_bounds__(128, 1) k_mma_multi(int iters, int nissue, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bar[4];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tptr;
  long long t0 = clock64();
  if (warp < nissue) {
    const uint32_t sb = smem_u32(smem);
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t bd0 = umma_desc_sw128(sb + 16384);
    const uint64_t ad0 = umma_desc_sw128(sb);
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t d = tbase + warp * N;
        if (elect_one()) tc_mma_f16(d, ad0 + 2 * (u & 3), bd0 + 2 * (u & 3), idesc, 1);
        __syncwarp();
      }
    }
    if (elect_one()) tc_commit(smem_u32(&bar[warp]));
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
  }
}

template <int N>
static void run_multi(long long *d_out, int nissue) {
  long long h[148];
  const int iters = 4000;
  cudaFuncSetAttribute(k_mma_multi<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k_mma_multi<N><<<148, 128, 64 * 1024>>>(iters, nissue, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("multi N=%d: %s\n", N, cudaGetErrorString(e));
      return;
    }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("mma 128x%dx16 SS, %d issuing warps: %.1f clk per instruction (all warps), floor %d\n", N, nissue,
         (double)mx / (iters * nissue), N / 2);
}

// One issuing warp (groups of 5 MMAs + commit, as the skinning phase of k_body_tc does) while
// `nld` other warps stream tcgen05.ld from other TMEM columns: does the epilogue's TMEM traffic
// slow the tensor pipe?
template <int N>
__global__ void __launch_bounds__(288, 1) k_mma_ld(int iters, int nld, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    done = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tptr;
  if (warp == 0) {
    const uint32_t sb = smem_u32(smem);
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t bd0 = umma_desc_sw128(sb + 16384);
    const uint64_t ad0 = umma_desc_sw128(sb);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 5) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 5; ++u) tc_mma_f16(tbase + ((i / 5) & 1) * N, ad0 + 2 * (u & 3), bd0 + 2 * (u & 3), idesc, u != 0);
        tc_commit(smem_u32(&bar[1]));
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(smem_u32(&bar[0]));
    __syncwarp();
    mbar_wait(smem_u32(&bar[0]), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    done = 1;
  } else if (warp - 1 < nld) {
    const uint32_t lb = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + 256;
    uint32_t acc = 0;
    while (!done) {
      uint32_t r[48];
      tc_ld_32x32(lb + (warp >> 2) * 64, r);
      tc_ld_32x16(lb + (warp >> 2) * 64 + 32, r + 32);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 48; ++j) acc ^= r[j];
    }
    if (acc == 0x12345u) out[1000] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
  }
}

template <int N>
static void run_mma_ld(long long *d_out, int nld) {
  long long h[148];
  const int iters = 4000;
  cudaFuncSetAttribute(k_mma_ld<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k_mma_ld<N><<<148, 288, 64 * 1024>>>(iters, nld, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mma_ld N=%d: %s\n", N, cudaGetErrorString(e));
      return;
    }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("mma 128x%dx16 SS in groups of 5 + commit, %d warps streaming tcgen05.ld: %.1f clk per MMA\n", N, nld, (double)mx / iters);
}

// tcgen05.ld throughput with several loads in flight per warp (K loads, then one wait::ld):
// is the cost per instruction or per byte?
template <int W>   // W = 2, 4, 8, 16, 32 columns per load
__global__ void __launch_bounds__(512, 1) k_ld_tp(int iters, long long *out) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tptr + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[4][32];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (W == 2) tc_ld_32x2(tbase + u * 32, r[u]);
      if (W == 4) tc_ld_32x4(tbase + u * 32, r[u]);
      if (W == 8) tc_ld_32x8(tbase + u * 32, r[u]);
      if (W == 16) tc_ld_32x16(tbase + u * 32, r[u]);
      if (W == 32) tc_ld_32x32(tbase + u * 32, r[u]);
    }
    tc_wait_ld();
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < W; ++j) acc ^= r[u][j];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) out[1000] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "n"(512) : "memory");
  }
}

template <int W>
static void run_ld_tp(long long *d_out, int warps) {
  long long h[148];
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) k_ld_tp<W><<<148, 32 * warps>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("ld_tp: %s\n", cudaGetErrorString(e));
    return;
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  double n_ld = (double)iters * 4 * warps;
  printf("tcgen05.ld x%-2d, %2d warps, 4 in flight each: %.1f clk per load per SM, %.1f B/clk/SM\n", W, warps, mx / n_ld,
         n_ld * 32 * W * 4 / mx);
}

int main() {
  long long *d_out;
  cudaMalloc(&d_out, 2048 * sizeof(long long));
  for (int w = 4; w <= 16; w *= 2) {
    run_ld_tp<2>(d_out, w);
    run_ld_tp<4>(d_out, w);
    run_ld_tp<8>(d_out, w);
    run_ld_tp<16>(d_out, w);
    run_ld_tp<32>(d_out, w);
  }
  for (int w = 0; w <= 8; w += 4) run_mma_ld<96>(d_out, w);
  run_mma_ld<192>(d_out, 0);
  run_mma_ld<192>(d_out, 8);
  for (int w = 1; w <= 1; ++w) run_multi<96>(d_out, w);
  for (int w = 1; w <= 4; w *= 2) run_multi<32>(d_out, w);
  long long h[148];
  for (int mode = 0; mode < 2; ++mode)
    for (int threads = 128; threads <= 256; threads += 128) {
      const int iters = 4000;
      for (int rep = 0; rep < 2; ++rep) k_ld<<<148, threads>>>(iters, mode, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("ld: %s\n", cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      double bytes = (double)iters * (threads / 32) * 32 * (mode == 0 ? 32 : 16) * 4;
      printf("tcgen05.ld x%d, %d warps: %.1f clk/iter, %.1f B/clk/SM\n", mode == 0 ? 32 : 16, threads / 32, (double)mx / iters,
             bytes / mx);
    }
  for (int nacc = 1; nacc <= 4; nacc *= 2) {
    run_mma<16>(d_out, 1, nacc);
    run_mma<16>(d_out, 0, nacc);
  }
  for (int nacc = 1; nacc <= 4; nacc *= 2) {
    run_mma<48>(d_out, 1, nacc);
    run_mma<48>(d_out, 0, nacc);
  }
  for (int nacc = 1; nacc <= 2; nacc *= 2) {
    run_mma<96>(d_out, 1, nacc);
    run_mma<96>(d_out, 0, nacc);
    run_mma<192>(d_out, 1, nacc);
    run_mma<192>(d_out, 0, nacc);
  }
  run_mma<128>(d_out, 0, 1);
  run_mma<128>(d_out, 0, 2);
  run_mma<64>(d_out, 0, 2);
  run_mma<64>(d_out, 1, 2);
  return 0;
}
