// Microbenchmark: verts [B][V][3] fp32 written through shared-memory staging + the TMA engine by ONE CTA of 8 warps per
// SM, a thread owning one vertex of 4 sample rows per step -- the epilogue of k_body_res -- in the kernel's walk order.
//   0: per warp, four cp.async.bulk of 384 B (one per sample row; 8-byte head / tail of odd rows as STG.64)
//   1: per 4 warps (one 128-vertex row segment), four cp.async.bulk of 1536 B, two named barriers per step
//   2: per warp, ONE cp.async.bulk.tensor box {96 floats, 4 rows}: rows of equal parity through a tensor map whose row
//      is TWO verts rows (stride 165,360 B is a multiple of 16; the odd rows sit at inner offset 3 V)
//   3: per 2 warps, one box {192 floats, 4 rows}
//   5: mode 0 without fence.proxy.async (timing only)
//   6: mode 2 without fence.proxy.async (timing only)
//   7: st.global reference (3 x STG.32 per vertex)
// Build: nvcc -arch=sm_100a -O3 --cudart=shared -o tma_store_pattern tma_store_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#define V 6890
#define NS 96
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float *out, const __grid_constant__ CUtensorMap map_e, const __grid_constant__ CUtensorMap map_o,
                                            int B, int n_vt, int n_m, int l2) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = n_vt * n_m;
  const int t0 = (int)((long long)blockIdx.x * total / gridDim.x), t1 = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
  const int q = warp & 3, part = warp >> 2;
  int rot = 0;
  const int n = t1 - t0;
  {
    const float step = n_m / 3.0f;
    for (int i = 0; i < n; ++i) {
      const int m = (t0 + i) % n_m;
      if (fmodf((float)m, step) < 1.0f) { rot = i; break; }
    }
  }
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const uint32_t sb = smem_u32(smem);
  for (int i = 0; i < n; ++i) {
    const int t = t0 + (i + rot) % n;
    const int vt = t / n_m, m = t % n_m;
    const int v0 = vt * 128 + 32 * q, v = v0 + lane;
    const int n_valid = min(32, V - v0);
    const bool ok = v < V;
    for (int st = 0; st < NS / 8; ++st) {
      // this warp's 4 samples of the 8-sample tile: modes 2, 3, 6 take equal parity, the others a contiguous half
      const bool par = MODE == 2 || MODE == 3 || MODE == 6 || MODE == 9;
      const int s_first = st * 8 + (par ? part : 4 * part), s_step = par ? 2 : 1;
      const int b0 = ((l2 & 1) ? 0 : m * NS) + s_first;
      float o[4][3];
#pragma unroll
      for (int r = 0; r < 4; ++r) { o[r][0] = __int_as_float((b0 + r * s_step) * 3 + v); o[r][1] = o[r][0] + 1.f; o[r][2] = o[r][0] + 2.f; }
      const int rows_left = (B - (m * NS + s_first) + s_step - 1) / s_step;
      if (MODE == 7) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r < rows_left && ok) {
            float *g = out + ((size_t)(b0 + r) * V + v) * 3;
            __stcs(g, o[r][0]); __stcs(g + 1, o[r][1]); __stcs(g + 2, o[r][2]);
          }
        continue;
      }
      if (MODE == 0 || MODE == 5) {
        const uint32_t stg = sb + warp * 1600;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        float *grow = out + ((size_t)b0 * V + v0) * 3;
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r < rows_left && ok) {
            const uint32_t head = (uint32_t)(uintptr_t)(grow + (size_t)r * V * 3) & 8u;
            const uint32_t d = stg + r * 400 + head + lane * 12;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(d), "f"(o[r][0]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 4), "f"(o[r][1]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 8), "f"(o[r][2]) : "memory");
          }
        if (MODE == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        const uint32_t seg = n_valid * 12;
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r < rows_left && n_valid > 0) {
            float *g = grow + (size_t)r * V * 3;
            const uint32_t head = (uint32_t)(uintptr_t)g & 8u, tail = (seg - head) & 8u, mid = seg - head - tail;
            if (lane == r && mid > 0)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"((char *)g + head),
                           "r"(stg + r * 400 + 2 * head), "r"(mid), "l"(pol)
                           : "memory");
            if (head && lane == 0) __stcs((float2 *)g, make_float2(o[r][0], o[r][1]));
            if (tail && lane == n_valid - 1) __stcs((float2 *)(g + 3 * n_valid - 2), make_float2(o[r][1], o[r][2]));
          }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      } else if (MODE == 1) {
        const uint32_t stg = sb + part * (4 * 1552);
        const int nv_tile = min(128, V - vt * 128);
        if (q == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
        float *grow = out + ((size_t)b0 * V + vt * 128) * 3;
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r < rows_left && ok) {
            const uint32_t head = (uint32_t)(uintptr_t)(grow + (size_t)r * V * 3) & 8u;
            const uint32_t d = stg + r * 1552 + head + (32 * q + lane) * 12;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(d), "f"(o[r][0]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 4), "f"(o[r][1]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 8), "f"(o[r][2]) : "memory");
          }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
        const uint32_t seg = nv_tile * 12;
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r < rows_left) {
            float *g = grow + (size_t)r * V * 3;
            const uint32_t head = (uint32_t)(uintptr_t)g & 8u, tail = (seg - head) & 8u, mid = seg - head - tail;
            if (q == 0 && lane == r && mid > 0)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"((char *)g + head),
                           "r"(stg + r * 1552 + 2 * head), "r"(mid), "l"(pol)
                           : "memory");
            if (head && q == 0 && lane == 0) __stcs((float2 *)g, make_float2(o[r][0], o[r][1]));
            if (tail && v == vt * 128 + nv_tile - 1) __stcs((float2 *)(g + 3 * nv_tile - 2), make_float2(o[r][1], o[r][2]));
          }
        if (q == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      } else if (MODE == 2 || MODE == 6) {
        const uint32_t stg = sb + warp * 1536;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint32_t d = stg + r * 384 + lane * 12;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(d), "f"(o[r][0]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 4), "f"(o[r][1]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 8), "f"(o[r][2]) : "memory");
        }
        if (MODE == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          // sample b0 = 2 y + parity: even rows through map_e (inner extent 3 V), odd rows through map_o at inner offset 3 V
          const int y = b0 >> 1;
          if ((b0 & 1) && !(l2 & 2))
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_o), "r"(stg),
                         "r"(3 * V + 3 * v0), "r"(y)
                         : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_e), "r"(stg),
                         "r"(3 * v0), "r"(y)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      } else if (MODE == 4) {
        // contiguous 4 samples (b0 even): rows 0, 2 through map_e as one box {96, 2}; rows 1, 3 start 8 bytes off the
        // 16-byte grid: box {92, 2} through map_o at inner offset 3 V + 3 v0 + 2, the first two and last two floats of
        // the segment as STG.64
        const uint32_t stg = sb + warp * 1536;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; r += 2) {
          const uint32_t d = stg + (r >> 1) * 384 + lane * 12;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(d), "f"(o[r][0]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 4), "f"(o[r][1]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 8), "f"(o[r][2]) : "memory");
        }
#pragma unroll
        for (int r = 1; r < 4; r += 2) {
          const uint32_t d = stg + 768 + (r >> 1) * 368 + lane * 12 - 8;
          if (lane > 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(d), "f"(o[r][0]) : "memory");
          if (lane > 0 && lane < 31) asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 4), "f"(o[r][1]) : "memory");
          if (lane < 31) asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 8), "f"(o[r][2]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        const int y = b0 >> 1;
        if (lane == 0)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_e), "r"(stg),
                       "r"(3 * v0), "r"(y)
                       : "memory");
        if (lane == 1)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_o), "r"(stg + 768),
                       "r"(3 * V + 3 * v0 + 2), "r"(y)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#pragma unroll
        for (int r = 1; r < 4; r += 2)
          if (r < rows_left && n_valid > 0) {
            float *g = out + ((size_t)(b0 + r) * V + v0) * 3;
            if (lane == 0) __stcs((float2 *)g, make_float2(o[r][0], o[r][1]));
            if (lane == 31 && n_valid == 32) __stcs((float2 *)(g + 94), make_float2(o[r][1], o[r][2]));
          }
      } else if (MODE == 9) {
        // equal-parity samples per warp (part = parity): even rows one box {96, 4}; odd rows one box {92, 4} at inner offset
        // 3 V + 3 v0 + 2 plus the first / last two floats of each row segment as STG.64
        const uint32_t stg = sb + warp * 1536;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (!(b0 & 1)) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const uint32_t d = stg + r * 384 + lane * 12;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(d), "f"(o[r][0]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 4), "f"(o[r][1]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 8), "f"(o[r][2]) : "memory");
          }
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const uint32_t d = stg + r * 368 + lane * 12 - 8;
            if (lane > 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(d), "f"(o[r][0]) : "memory");
            if (lane > 0 && lane < 31) asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 4), "f"(o[r][1]) : "memory");
            if (lane < 31) asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 8), "f"(o[r][2]) : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        const int y = b0 >> 1;
        if (lane == 0) {
          if (b0 & 1)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_o), "r"(stg),
                         "r"(3 * V + 3 * v0 + 2), "r"(y)
                         : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_e), "r"(stg),
                         "r"(3 * v0), "r"(y)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (b0 & 1) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r < rows_left && n_valid > 0) {
              float *g = out + ((size_t)(b0 + 2 * r) * V + v0) * 3;
              if (lane == 0) __stcs((float2 *)g, make_float2(o[r][0], o[r][1]));
              if (lane == 31 && n_valid == 32) __stcs((float2 *)(g + 94), make_float2(o[r][1], o[r][2]));
            }
        }
      } else if (MODE == 3) {
        // two warps (quarters 2 h, 2 h + 1 of one parity) share a box of 192 floats x 4 rows
        const int h = q >> 1;
        const uint32_t stg = sb + (part * 2 + h) * 3072;
        const int bar = 1 + part * 2 + h;
        if ((q & 1) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint32_t d = stg + r * 768 + ((q & 1) * 32 + lane) * 12;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(d), "f"(o[r][0]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 4), "f"(o[r][1]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(d + 8), "f"(o[r][2]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
        if ((q & 1) == 0 && lane == 0) {
          const int y = b0 >> 1, x = 3 * (vt * 128 + 64 * h);
          if (b0 & 1)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_o), "r"(stg),
                         "r"(3 * V + x), "r"(y)
                         : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map_e), "r"(stg), "r"(x),
                         "r"(y)
                         : "memory");
        }
        if ((q & 1) == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                          CUtensorMapFloatOOBfill);
static enc_t g_enc;
static void make_maps(CUtensorMap *me, CUtensorMap *mo, float *buf, int B, int box_inner) {
  const int mode4 = box_inner == 0, mode9 = box_inner == -1;
  if (mode9) box_inner = 96;
  cuuint64_t strides[1] = {(cuuint64_t)V * 24};
  cuuint32_t box[2] = {(cuuint32_t)(mode4 ? 96 : box_inner), (cuuint32_t)(mode4 ? 2 : 4)}, estr[2] = {1, 1};
  cuuint64_t de[2] = {(cuuint64_t)3 * V, (cuuint64_t)(B + 1) / 2}, dod[2] = {(cuuint64_t)6 * V, (cuuint64_t)B / 2};
  CUresult r1 = g_enc(me, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, de, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (mode4 || mode9) box[0] = 92;
  CUresult r2 = g_enc(mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dod, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 || r2) printf("cuTensorMapEncodeTiled failed: %d %d\n", (int)r1, (int)r2);
}

static float *g_ref;
template <int MODE>
void run(const char *name, float *buf, int B, int l2, int box_inner, bool check) {
  int n_vt = (V + 127) / 128, n_m = (B + NS - 1) / NS;
  alignas(64) CUtensorMap me, mo;
  make_maps(&me, &mo, buf, B, box_inner);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  if (check) cudaMemset(buf, 0xff, (size_t)B * V * 12);
  for (int it = 0; it < 3; ++it) k<MODE><<<148, 256, 16384>>>(buf, me, mo, B, n_vt, n_m, l2);
  cudaEventRecord(e0);
  for (int it = 0; it < 10; ++it) k<MODE><<<148, 256, 16384>>>(buf, me, mo, B, n_vt, n_m, l2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
  long long bad = -1;
  if (check && !l2) {
    // every element must equal the st.global reference
    static float *h = nullptr, *hr = nullptr;
    size_t nel = (size_t)B * V * 3;
    if (!h) { h = (float *)malloc(nel * 4); hr = (float *)malloc(nel * 4); cudaMemcpy(hr, g_ref, nel * 4, cudaMemcpyDeviceToHost); }
    cudaMemcpy(h, buf, nel * 4, cudaMemcpyDeviceToHost);
    bad = 0;
    for (size_t i = 0; i < nel; ++i) bad += memcmp(&h[i], &hr[i], 4) != 0;
  }
  printf("%s%-44s: %.1f us -> %.0f GB/s  mismatches %lld (%s)\n", l2 ? "[L2-resident] " : "", name, ms * 1e3,
         (double)B * V * 12 / ms / 1e6, bad, cudaGetErrorString(cudaGetLastError()));
}
int main(int argc, char **argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  g_enc = (enc_t)fn;
  const int B = 4096;
  float *buf;
  cudaMalloc(&buf, (size_t)B * V * 12 + 256);
  cudaMalloc(&g_ref, (size_t)B * V * 12 + 256);
  run<7>("7 st.global 3 x STG.32 (reference)", g_ref, B, 0, 96, false);
  if (only == 2) { run<2>("2 tensor box {96, 4} per warp", buf, B, 0, 96, true); return 0; }
  if (only == 20) { run<2>("2 tensor box {96, 4} per warp, even map only", buf, B, 2, 96, false); return 0; }
  if (only == 4) { run<4>("4 tensor boxes {96, 2} + {92, 2} + head / tail STG.64", buf, B, 0, 0, true); return 0; }
  if (only == 9) { run<9>("9 parity split: box {96, 4} / box {92, 4} + head / tail STG.64", buf, B, 0, -1, true); return 0; }
  if (only == 3) { run<3>("3 tensor box {192, 4} per 2 warps", buf, B, 0, 192, true); return 0; }
  for (int l2 : {0, 1}) {
    run<7>("7 st.global 3 x STG.32", buf, B, l2, 96, true);
    run<0>("0 bulk 384 B per warp-row", buf, B, l2, 96, true);
    run<5>("5 ... without fence.proxy.async", buf, B, l2, 96, false);
    run<1>("1 bulk 1536 B per 4-warp row", buf, B, l2, 96, true);
  }
  return 0;
}
