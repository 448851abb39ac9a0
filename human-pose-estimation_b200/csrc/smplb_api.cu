// C ABI of libsmplb.so (see include/smplb.h): context, memory, orchestration of the kernels.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>

#include "smplb_internal.h"

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";

void smplb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *smplb_last_error(void) { return g_err; }
extern "C" int smplb_version(void) { return SMPLB_VERSION; }

// ---------------------------------------------------------------------------------- profiling
ProfScope::ProfScope(smplb_ctx *ctx, const char *name) : c(ctx) {
  r.name = name;
  r.e0 = r.e1 = nullptr;
  if (!c->profile) return;
  for (cudaEvent_t *e : {&r.e0, &r.e1}) {
    if (!c->event_pool.empty()) {
      *e = c->event_pool.back();
      c->event_pool.pop_back();
    } else {
      cudaEventCreate(e);
    }
  }
  cudaEventRecord(r.e0, c->cur);
}
ProfScope::~ProfScope() {
  if (!r.e0) return;
  cudaEventRecord(r.e1, c->cur);
  c->prof.push_back(r);
}

// ------------------------------------------------------------------------- host-mode staging
// In SMPLB_HOST mode every pointer is host memory: inputs are copied to temporary device
// buffers (stream-ordered allocations), outputs are copied back at finish().
struct Stager {
  smplb_ctx *c;
  int mem;
  struct Out {
    void *host, *dev;
    size_t bytes;
  };
  std::vector<void *> temps;
  std::vector<Out> outs;
  bool failed = false;
  bool async = false;
  Stager(smplb_ctx *ctx, int m) : c(ctx), mem(m == SMPLB_HOST_ASYNC ? SMPLB_HOST : m), async(m == SMPLB_HOST_ASYNC) {}
  void *alloc(size_t bytes) {
    void *d = nullptr;
    if (cudaMallocAsync(&d, bytes ? bytes : 4, c->stream) != cudaSuccess) {
      failed = true;
      return nullptr;
    }
    temps.push_back(d);
    return d;
  }
  template <typename T>
  const T *in(const T *p, size_t n) {
    if (!p || mem == SMPLB_DEVICE) return p;
    void *d = alloc(n * sizeof(T));
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) failed = true;
    return (const T *)d;
  }
  template <typename T>
  T *out(T *p, size_t n) {
    if (!p || mem == SMPLB_DEVICE) return p;
    void *d = alloc(n * sizeof(T));
    if (!d) return nullptr;
    outs.push_back({(void *)p, d, n * sizeof(T)});
    return (T *)d;
  }
  int finish() {
    int rc = 0;
    if (mem == SMPLB_HOST) {
      for (auto &o : outs)
        if (cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) failed = true;
      for (void *t : temps) cudaFreeAsync(t, c->stream);
      temps.clear();
      cudaError_t e = async ? cudaSuccess : cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess || failed) {
        smplb_set_error("host-mode staging failed: %s", cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
        rc = SMPLB_ECUDA;
      }
    }
    return rc;
  }
  ~Stager() {
    for (void *t : temps) cudaFreeAsync(t, c->stream);
  }
};

#define CHECK_CTX(c)                                                    \
  do {                                                                  \
    RET_IF(!(c), SMPLB_EINVAL, "null context");                         \
    CUDA_TRY(cudaSetDevice((c)->device));                               \
    (c)->cur = (c)->stream;                                             \
  } while (0)
#define CHECK_MEM(mem)                                                                                   \
  RET_IF((mem) != SMPLB_HOST && (mem) != SMPLB_DEVICE && (mem) != SMPLB_HOST_ASYNC, SMPLB_EINVAL,         \
         "mem must be SMPLB_HOST, SMPLB_DEVICE or SMPLB_HOST_ASYNC")

// ------------------------------------------------------------------------------- create-time
// J0 = J_regressor^T v_template and Jdirs = J_regressor^T shapedirs in fp64 (one-off constant
// folding of batch_smpl.py:115-118 through :110-112; exact algebra).
__global__ void __launch_bounds__(256) k_fold_joints(int V, int NB, const float *__restrict__ Jreg,
                                                     const float *__restrict__ shapedirs, const float *__restrict__ vt,
                                                     float *__restrict__ J0, float *__restrict__ Jdirs) {
  __shared__ double red[256];
  int jc = blockIdx.x;           // 3j + c
  int k = blockIdx.y;            // 0..NB-1: Jdirs column, NB: J0
  int j = jc / 3, cc = jc % 3;
  double s = 0.0;
  for (int v = threadIdx.x; v < V; v += 256) {
    double src = (k < NB) ? (double)shapedirs[(size_t)k * V * 3 + 3 * v + cc] : (double)vt[3 * v + cc];
    s += (double)Jreg[(size_t)v * NJ + j] * src;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (k < NB) Jdirs[(size_t)jc * NB + k] = (float)red[0];
    else J0[jc] = (float)red[0];
  }
}

// Dext[k][c * Vp + v] (planar, zero padded): rows 0..206 posedirs, 207..207+NB-1 shapedirs,
// 207+NB v_template -- a pure re-layout of the reference's [*, 3v+c] matrices.
__global__ void k_build_dext(int V, int Vp, int NB, const float *__restrict__ posedirs,
                             const float *__restrict__ shapedirs, const float *__restrict__ vt, float *__restrict__ Dext) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int k = blockIdx.y;
  if (n >= 3 * Vp) return;
  int cc = n / Vp, v = n % Vp;
  float val = 0.f;
  if (v < V) {
    size_t src = 3 * (size_t)v + cc;
    if (k < NPF) val = posedirs[(size_t)k * 3 * V + src];
    else if (k < NPF + NB) val = shapedirs[(size_t)(k - NPF) * 3 * V + src];
    else if (k == NPF + NB) val = vt[src];
  }
  Dext[(size_t)k * 3 * Vp + n] = val;
}

// Dext_act[k][c * Vpa + a] = Dext[k][c * Vp + act[a]]
__global__ void k_gather_dext(int n_act, int Vpa, int Vp, const int *__restrict__ act, const float *__restrict__ Dext,
                              float *__restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int k = blockIdx.y;
  if (n >= 3 * Vpa) return;
  int cc = n / Vpa, a = n % Vpa;
  out[(size_t)k * 3 * Vpa + n] = a < n_act ? Dext[(size_t)k * 3 * Vp + (size_t)cc * Vp + act[a]] : 0.f;
}

template <typename T>
static int dev_upload(T **dst, const T *src, size_t n) {
  CUDA_TRY(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) CUDA_TRY(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

static void free_ws(smplb_ctx *c) {
  void **ptrs[] = {(void **)&c->ws_x, (void **)&c->ws_Rs, (void **)&c->ws_J, (void **)&c->ws_A, (void **)&c->ws_Jtr,
                   (void **)&c->ws_vposed, (void **)&c->ws_verts, (void **)&c->ws_joints, (void **)&c->ws_kp,
                   (void **)&c->ws_dkp, (void **)&c->ws_djoints, (void **)&c->ws_dverts, (void **)&c->ws_dp, (void **)&c->ws_dA,
                   (void **)&c->ws_dx, (void **)&c->ws_part, (void **)&c->ws_cnt, (void **)&c->ws_theta,
                   (void **)&c->ws_beta, (void **)&c->ws_gp, (void **)&c->ws_x16, (void **)&c->ws_dp_act, (void **)&c->ws_A16, (void **)&c->ws_vposed_act, (void **)&c->ws_verts_act, (void **)&c->ws_U, (void **)&c->ws_du16,
                   (void **)&c->ws_rowscale, (void **)&c->ws_x16b, (void **)&c->ws_dp16};
  for (void **p : ptrs) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  c->ws_batch = 0;
  c->saved_B = 0;
  c->ws_gp_cap = 0;
}

#define WS_ALLOC(field, count)                                                        \
  do {                                                                                \
    size_t _n = (size_t)(count);                                                      \
    CUDA_TRY(cudaMalloc((void **)&c->field, std::max<size_t>(_n, 1) * 4));            \
  } while (0)

// Big buffers that only some calls need are allocated lazily by ensure_* below.
static int ensure_ws(smplb_ctx *c, int B) {
  if (B <= c->ws_batch) return 0;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  free_ws(c);
  int n = std::max(B, c->max_batch);
  size_t nb = (size_t)n;
  WS_ALLOC(ws_x, nb * KX);
  WS_ALLOC(ws_x16, nb * 128);   // 256 halves per row
  if (c->fold_nup) {
    WS_ALLOC(ws_U, nb * c->fold_nup);
    WS_ALLOC(ws_du16, nb * 3 * c->fold_nup / 2);
    CUDA_TRY(cudaMemsetAsync(c->ws_du16, 0, nb * 3 * c->fold_nup * 2, c->stream));
    WS_ALLOC(ws_rowscale, nb);
  }
  WS_ALLOC(ws_A16, nb * 12 * 32);   // 12 rows x 64 halves per sample; columns 48..63 stay zero
  CUDA_TRY(cudaMemsetAsync(c->ws_A16, 0, nb * 12 * 32 * 4, c->stream));
  WS_ALLOC(ws_Rs, nb * NJ * 9);
  WS_ALLOC(ws_J, nb * NJ * 3);
  WS_ALLOC(ws_A, nb * NJ * 12);
  WS_ALLOC(ws_Jtr, nb * NJ * 3);
  WS_ALLOC(ws_vposed, nb * c->pitch);
  WS_ALLOC(ws_joints, nb * c->K * 3);
  WS_ALLOC(ws_kp, nb * c->K * 2);
  WS_ALLOC(ws_dkp, nb * c->K * 2);
  WS_ALLOC(ws_djoints, nb * c->K * 3);
  WS_ALLOC(ws_dA, skin_bwd_part_rows((int)nb) * NJ * 12);
  WS_ALLOC(ws_dx, (size_t)c->ksplit * (nb + 128) * KX);   // split-K partials: up to 16 x (B rounded up to 128) rows
  WS_ALLOC(ws_part, nb);
  WS_ALLOC(ws_cnt, nb);
  WS_ALLOC(ws_theta, nb * 72);
  WS_ALLOC(ws_beta, nb * c->NB);
  c->ws_batch = n;
  return 0;
}

static int ensure_buf(smplb_ctx *c, float **field, size_t count, bool zero) {
  if (*field) return 0;
  CUDA_TRY(cudaMalloc((void **)field, std::max<size_t>(count, 1) * 4));
  if (zero) CUDA_TRY(cudaMemsetAsync(*field, 0, count * 4, c->stream));
  return 0;
}

extern "C" int smplb_create(smplb_ctx **out, const smplb_model *m, int device, int max_batch) {
  RET_IF(!out || !m, SMPLB_EINVAL, "null argument");
  *out = nullptr;
  RET_IF(m->num_verts < 1 || m->num_betas < 1 || m->num_betas > KX - NPF - 1, SMPLB_EINVAL,
         "num_verts >= 1 and 1 <= num_betas <= %d required", KX - NPF - 1);
  RET_IF(m->num_keypoints < 1 || m->num_keypoints > MAXK, SMPLB_EINVAL, "1 <= num_keypoints <= %d required", MAXK);
  RET_IF(!m->v_template || !m->shapedirs || !m->posedirs || !m->J_regressor || !m->weights || !m->joint_regressor ||
             !m->parents,
         SMPLB_EINVAL, "model has a null array");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  RET_IF(e != cudaSuccess || ndev == 0, SMPLB_EDEVICE, "no CUDA device available (%s); libsmplb has no CPU fallback",
         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  RET_IF(device < 0 || device >= ndev, SMPLB_EINVAL, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  RET_IF(prop.major != 10, SMPLB_EDEVICE, "device %d is sm_%d%d; libsmplb is built for sm_100a only", device, prop.major,
         prop.minor);
  CUDA_TRY(cudaSetDevice(device));

  smplb_ctx *c = new smplb_ctx();
  c->device = device;
  c->V = m->num_verts;
  c->NB = m->num_betas;
  c->K = m->num_keypoints;
  c->V3 = 3 * c->V;
  c->Vp = cdiv(c->V, 128) * 128;
  c->pitch = 3 * c->Vp;
  c->ksplit = 16;
  c->max_batch = std::max(max_batch, 1);
  if (const char *e = getenv("SMPLB_L2_CHUNK")) c->l2_chunk = atoi(e);   // experiment knob, see DESIGN.md section 4
  // kinematic tree: parent index < child index (batch_lbs.py:130 walks joints in index order)
  c->tree.max_depth = 0;
  for (int j = 0; j < NJ; ++j) {
    int p = m->parents[j];
    if (p < 0 || p >= NJ) p = -1;    // uint32(-1) cast to int32 (batch_smpl.py:65)
    if (j == 0) p = -1;
    if (p >= j) {
      delete c;
      smplb_set_error("parents[%d] = %d: parents must precede children", j, p);
      return SMPLB_EINVAL;
    }
    c->tree.parent[j] = (signed char)p;
    c->tree.depth[j] = p < 0 ? 0 : (signed char)(c->tree.depth[p] + 1);
    c->tree.max_depth = std::max(c->tree.max_depth, (int)c->tree.depth[j]);
  }
  {
    int nch[NJ] = {};
    for (int j = 0; j < NJ; ++j) c->tree.child[j][0] = c->tree.child[j][1] = c->tree.child[j][2] = -1;
    for (int j = 0; j < NJ; ++j) {
      const int p = c->tree.parent[j];
      if (p < 0) continue;
      if (nch[p] < 3) c->tree.child[p][nch[p]] = (signed char)j;
      nch[p]++;
    }
    c->tree.level_slots = 0;
    c->tree.max_children = 0;
    for (int j = 0; j < NJ; ++j) {
      c->tree.max_children = std::max(c->tree.max_children, nch[j]);
      const int d = c->tree.depth[j] + 1;
      if (d < 16) {
        const unsigned cur = (c->tree.level_slots >> (2 * d)) & 3u;
        const unsigned want = (unsigned)std::min(nch[j], 3);
        if (want > cur) c->tree.level_slots = (c->tree.level_slots & ~(3u << (2 * d))) | (want << (2 * d));
      }
    }
  }
  int rc = 0;
  auto fail = [&](int code) {
    smplb_destroy(c);
    return code;
  };
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);   // (equal when the device has no priorities)
  // the vertex kernel and the GEMMs on low-priority streams, the per-body kernels on high-priority ones (DESIGN.md section 4;
  // vertex kernel high / all equal were measured too: 162.0-162.9 vs 162.9-164.3 us per step, no difference)
  if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
      cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_least) != cudaSuccess ||
      cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
      cudaStreamCreateWithPriority(&c->stream_g, cudaStreamNonBlocking, prio_least) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_g_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_g_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_red_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_red_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_step0, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_cnt, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess)
    return fail(SMPLB_ECUDA);
  c->cur = c->stream;
  {
    // host-mode calls stage through stream-ordered allocations: keep the pool's memory across
    // synchronisations instead of returning it to the driver after every call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  size_t V = c->V;
  if ((rc = dev_upload(&c->d_vt, m->v_template, V * 3))) return fail(rc);
  if ((rc = dev_upload(&c->d_shapedirs, m->shapedirs, (size_t)c->NB * V * 3))) return fail(rc);
  if ((rc = dev_upload(&c->d_posedirs, m->posedirs, (size_t)NPF * V * 3))) return fail(rc);
  if ((rc = dev_upload(&c->d_W, m->weights, V * NJ))) return fail(rc);
  if ((rc = dev_upload(&c->d_JR, m->joint_regressor, V * c->K))) return fail(rc);
  float *d_Jreg = nullptr;
  if ((rc = dev_upload(&d_Jreg, m->J_regressor, V * NJ))) return fail(rc);
  // Dext = [posedirs ; shapedirs ; v_template ; 0], planar columns, row pitch c->pitch
  size_t dext_bytes = (size_t)KX * c->pitch * sizeof(float);
  if (cudaMalloc((void **)&c->d_Dext, dext_bytes) != cudaSuccess) return fail(SMPLB_ECUDA);
  k_build_dext<<<dim3(cdiv(c->pitch, 256), KX), 256, 0, c->stream>>>(c->V, c->Vp, c->NB, c->d_posedirs, c->d_shapedirs,
                                                                     c->d_vt, c->d_Dext);
  c->launches++;
  if (cudaMalloc((void **)&c->d_J0, NJ * 3 * 4) != cudaSuccess ||
      cudaMalloc((void **)&c->d_Jdirs, (size_t)NJ * 3 * c->NB * 4) != cudaSuccess)
    return fail(SMPLB_ECUDA);
  k_fold_joints<<<dim3(NJ * 3, c->NB + 1), 256, 0, c->stream>>>(c->V, c->NB, d_Jreg, c->d_shapedirs, c->d_vt, c->d_J0,
                                                               c->d_Jdirs);
  c->launches++;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
    smplb_set_error("k_fold_joints failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_Jreg);
    return fail(SMPLB_ECUDA);
  }
  cudaFree(d_Jreg);
  // CSR views of joint_regressor [V,K] (index construction only, any sparsity pattern)
  {
    int K = c->K;
    std::vector<int> koff(K + 1, 0), kidx, voff(V + 1, 0), vk;
    std::vector<float> kval, vval;
    for (int k = 0; k < K; ++k) {
      for (size_t v = 0; v < V; ++v) {
        float w = m->joint_regressor[v * K + k];
        if (w != 0.0f) {
          kidx.push_back((int)v);
          kval.push_back(w);
        }
      }
      koff[k + 1] = (int)kidx.size();
    }
    for (size_t v = 0; v < V; ++v) {
      for (int k = 0; k < K; ++k) {
        float w = m->joint_regressor[v * K + k];
        if (w != 0.0f) {
          vk.push_back(k);
          vval.push_back(w);
        }
      }
      voff[v + 1] = (int)vk.size();
    }
    if ((rc = dev_upload(&c->d_kcsr_off, koff.data(), koff.size()))) return fail(rc);
    if ((rc = dev_upload(&c->d_kcsr_idx, kidx.data(), kidx.size()))) return fail(rc);
    if ((rc = dev_upload(&c->d_kcsr_val, kval.data(), kval.size()))) return fail(rc);
    if ((rc = dev_upload(&c->d_vcsr_off, voff.data(), voff.size()))) return fail(rc);
    if ((rc = dev_upload(&c->d_vcsr_k, vk.data(), vk.size()))) return fail(rc);
    if ((rc = dev_upload(&c->d_vcsr_val, vval.data(), vval.size()))) return fail(rc);
    // active vertices (non-empty rows) and everything the compact backward needs, gathered
    std::vector<int> act, aoff(1, 0), ak;
    std::vector<float> aval, aW;
    for (size_t v = 0; v < V; ++v) {
      if (voff[v + 1] == voff[v]) continue;
      act.push_back((int)v);
      for (int e = voff[v]; e < voff[v + 1]; ++e) {
        ak.push_back(vk[e]);
        aval.push_back(vval[e]);
      }
      aoff.push_back((int)ak.size());
      for (int j = 0; j < NJ; ++j) aW.push_back(m->weights[v * NJ + j]);
    }
    c->n_act = (int)act.size();
    {
      std::vector<int> slot_of(V, -1), kslot(kidx.size());
      for (size_t a = 0; a < act.size(); ++a) slot_of[act[a]] = (int)a;
      for (size_t e = 0; e < kidx.size(); ++e) kslot[e] = slot_of[kidx[e]];
      if ((rc = dev_upload(&c->d_kcsr_slot, kslot.data(), kslot.size()))) return fail(rc);
    }
    c->Vpa = std::max(128, cdiv(c->n_act, 128) * 128);
    c->pitch_act = 3 * c->Vpa;
    if ((rc = dev_upload(&c->d_act_idx, act.data(), act.size()))) return fail(rc);
    if ((rc = dev_upload(&c->d_act_W, aW.data(), aW.size()))) return fail(rc);
    if ((rc = dev_upload(&c->d_acsr_off, aoff.data(), aoff.size()))) return fail(rc);
    if ((rc = dev_upload(&c->d_acsr_k, ak.data(), ak.size()))) return fail(rc);
    if ((rc = dev_upload(&c->d_acsr_val, aval.data(), aval.size()))) return fail(rc);
    size_t ab = (size_t)KX * c->pitch_act * sizeof(float);
    if (cudaMalloc((void **)&c->d_Dext_act, ab) != cudaSuccess) return fail(SMPLB_ECUDA);
    k_gather_dext<<<dim3(cdiv(c->pitch_act, 256), KX), 256, 0, c->stream>>>(c->n_act, c->Vpa, c->Vp, c->d_act_idx,
                                                                           c->d_Dext, c->d_Dext_act);
    c->launches++;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail(SMPLB_ECUDA);
  }
  if (cudaMalloc((void **)&c->ws_scal, 64 * 4) != cudaSuccess ||
      cudaMalloc((void **)&c->ws_cnt64, 8 * sizeof(long long)) != cudaSuccess ||
      cudaMalloc((void **)&c->x_status, sizeof(int)) != cudaSuccess)
    return fail(SMPLB_ECUDA);
  cudaMemset(c->x_status, 0, sizeof(int));
  cudaMemset(c->ws_scal, 0, 64 * 4);
  cudaMemset(c->ws_cnt64, 0, 8 * sizeof(long long));
  if ((rc = blend_tc_init(c))) return fail(rc);
  if ((rc = skin_tc_init(c))) return fail(rc);
  if ((rc = skin_bwd_tc_init(c))) return fail(rc);
  if ((rc = compact_tc_init(c))) return fail(rc);
  if ((rc = body_tc_init(c))) return fail(rc);
  if ((rc = fold_init(c))) return fail(rc);
  if ((rc = ensure_ws(c, c->max_batch))) return fail(rc);
  *out = c;
  return 0;
}

extern "C" int smplb_destroy(smplb_ctx *c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  smplb_comm_destroy(c);
  free_ws(c);
  void *ptrs[] = {c->d_vt,       c->d_shapedirs, c->d_posedirs, c->d_W,        c->d_JR,       c->d_Dext,   c->d_J0,
                  c->d_Jdirs,    c->d_kcsr_off,  c->d_kcsr_idx, c->d_kcsr_val, c->d_vcsr_off, c->d_vcsr_k, c->d_vcsr_val,
                  c->ws_scal,    c->ws_cnt64,    c->flush_buf,   c->d_Dt16,     c->d_W16,      c->d_WT16,      c->d_G,        c->d_cc,       c->d_G16,      c->d_Gt16,     c->d_Dt16_act, c->d_W16_act,  c->d_kcsr_slot, c->d_act_idx,  c->d_act_W,  c->d_acsr_off,
                  c->d_acsr_k,   c->d_acsr_val,  c->d_Dext_act,   c->ws_silpred, c->ws_dsil,    c->ws_silcnt, c->ws_mesh_part, c->ws_grid,  c->ws_vdist, c->ws_segpts, c->ws_segoff, c->ws_segbits, c->ws_ticket,
                  c->x_mbox,     c->x_status,    c->d_Dbf};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (int i = 0; i < 16; ++i) {
    if (c->timer0[i]) cudaEventDestroy(c->timer0[i]);
    if (c->timer1[i]) cudaEventDestroy(c->timer1[i]);
  }
  for (auto &r : c->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto e : c->event_pool) cudaEventDestroy(e);
  if (c->stream2) {
    cudaStreamSynchronize(c->stream2);
    cudaStreamDestroy(c->stream2);
  }
  if (c->stream3) {
    cudaStreamSynchronize(c->stream3);
    cudaStreamDestroy(c->stream3);
  }
  if (c->stream_g) {
    cudaStreamSynchronize(c->stream_g);
    cudaStreamDestroy(c->stream_g);
  }
  if (c->ev_g_fork) cudaEventDestroy(c->ev_g_fork);
  if (c->ev_g_join) cudaEventDestroy(c->ev_g_join);
  if (c->ev_red_fork) cudaEventDestroy(c->ev_red_fork);
  if (c->ev_red_join) cudaEventDestroy(c->ev_red_join);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_step0) cudaEventDestroy(c->ev_step0);
  if (c->ev_cnt) cudaEventDestroy(c->ev_cnt);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

// ------------------------------------------------------------------------------------- memory
extern "C" int smplb_malloc(smplb_ctx *c, void **dptr, size_t bytes) {
  CHECK_CTX(c);
  RET_IF(!dptr, SMPLB_EINVAL, "null dptr");
  // stream-ordered pool allocation: no device synchronisation, memory is reused across calls
  CUDA_TRY(cudaMallocAsync(dptr, bytes ? bytes : 1, c->stream));
  return 0;
}
extern "C" int smplb_free(smplb_ctx *c, void *dptr) {
  CHECK_CTX(c);
  if (dptr) CUDA_TRY(cudaFreeAsync(dptr, c->stream));   // ordered after every use on the context's stream
  return 0;
}
extern "C" int smplb_host_alloc(void **hptr, size_t bytes) {
  RET_IF(!hptr, SMPLB_EINVAL, "null hptr");
  CUDA_TRY(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
extern "C" int smplb_host_free(void *hptr) {
  CUDA_TRY(cudaFreeHost(hptr));
  return 0;
}
extern "C" int smplb_memcpy_h2d(smplb_ctx *c, void *dst, const void *src, size_t bytes) {
  CHECK_CTX(c);
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return 0;
}
extern "C" int smplb_memcpy_d2h(smplb_ctx *c, void *dst, const void *src, size_t bytes) {
  CHECK_CTX(c);
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}
extern "C" int smplb_memset(smplb_ctx *c, void *dst, int value, size_t bytes) {
  CHECK_CTX(c);
  CUDA_TRY(cudaMemsetAsync(dst, value, bytes, c->stream));
  return 0;
}
extern "C" int smplb_sync(smplb_ctx *c) {
  CHECK_CTX(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int smplb_order_after(smplb_ctx *c, smplb_ctx *other) {
  CHECK_CTX(c);
  RET_IF(!other, SMPLB_EINVAL, "null context");
  RET_IF(other->device != c->device, SMPLB_EINVAL, "contexts live on different devices");
  cudaEvent_t ev;
  CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(ev, other->stream));
  CUDA_TRY(cudaStreamWaitEvent(c->stream, ev, 0));
  CUDA_TRY(cudaEventDestroy(ev));   // released once the wait has been satisfied
  return 0;
}
extern "C" int smplb_flush_l2(smplb_ctx *c, size_t bytes) {
  CHECK_CTX(c);
  if (bytes > c->flush_bytes) {
    if (c->flush_buf) CUDA_TRY(cudaFree(c->flush_buf));
    c->flush_buf = nullptr;
    CUDA_TRY(cudaMalloc(&c->flush_buf, bytes));
    c->flush_bytes = bytes;
  }
  CUDA_TRY(cudaMemsetAsync(c->flush_buf, 0x5a, bytes, c->stream));
  return 0;
}
extern "C" int smplb_timer_start(smplb_ctx *c, int slot) {
  CHECK_CTX(c);
  RET_IF(slot < 0 || slot >= 16, SMPLB_EINVAL, "timer slot out of range");
  if (!c->timer0[slot]) {
    CUDA_TRY(cudaEventCreate(&c->timer0[slot]));
    CUDA_TRY(cudaEventCreate(&c->timer1[slot]));
  }
  CUDA_TRY(cudaEventRecord(c->timer0[slot], c->stream));
  return 0;
}
extern "C" int smplb_timer_stop(smplb_ctx *c, int slot) {
  CHECK_CTX(c);
  RET_IF(slot < 0 || slot >= 16 || !c->timer1[slot], SMPLB_EINVAL, "timer slot not started");
  CUDA_TRY(cudaEventRecord(c->timer1[slot], c->stream));
  return 0;
}
extern "C" int smplb_timer_elapsed_ms(smplb_ctx *c, int slot, float *ms) {
  CHECK_CTX(c);
  RET_IF(slot < 0 || slot >= 16 || !c->timer1[slot] || !ms, SMPLB_EINVAL, "timer slot not started");
  CUDA_TRY(cudaEventSynchronize(c->timer1[slot]));
  CUDA_TRY(cudaEventElapsedTime(ms, c->timer0[slot], c->timer1[slot]));
  return 0;
}
extern "C" int smplb_launch_count(smplb_ctx *c, int64_t *count) {
  RET_IF(!c || !count, SMPLB_EINVAL, "null argument");
  *count = c->launches;
  return 0;
}
extern "C" int smplb_debug_set(smplb_ctx *c, const char *key, int value) {
  RET_IF(!c || !key, SMPLB_EINVAL, "null argument");
  if (!strcmp(key, "blend_tc")) {
    c->use_tc = value;
    return 0;
  }
  if (!strcmp(key, "blend_bwd_tc")) {
    c->use_blend_bwd_tc = value;
    return 0;
  }
  if (!strcmp(key, "comm_backend")) {
    c->comm_backend = value;
    return 0;
  }
  if (!strcmp(key, "comm_timeout_ms")) {
    c->x_timeout_ns = (unsigned long long)(value > 0 ? value : 1) * 1000000ull;
    return 0;
  }
  if (!strcmp(key, "mesh_grid")) {
    c->use_mesh_grid = value;
    return 0;
  }
  if (!strcmp(key, "fold_warp")) {
    c->fold_warp_kernels = value;
    return 0;
  }
  if (!strcmp(key, "overlap")) {
    c->use_overlap = value;
    return 0;
  }
  if (!strcmp(key, "fold")) {
    c->use_fold = value;
    return 0;
  }
  if (!strcmp(key, "l2_chunk")) {
    c->l2_chunk = value;
    return 0;
  }
  if (!strcmp(key, "fold_step")) {
    c->use_fold_step = value;
    return 0;
  }
  if (!strcmp(key, "fused")) {
    c->use_fused = value;
    return 0;
  }
  if (!strcmp(key, "prio")) {
    c->use_prio = value;
    return 0;
  }
  if (!strcmp(key, "body_pairs")) {
    c->body_pairs = value;
    return 0;
  }
  if (!strcmp(key, "skin_tc")) {
    c->use_skin_tc = value;
    return 0;
  }
  if (!strcmp(key, "skin_bwd_tc")) {
    c->use_skin_bwd_tc = value;
    return 0;
  }
  if (!strcmp(key, "mesh_lattice")) {
    c->use_mesh_lattice = value;
    return 0;
  }
  if (!strcmp(key, "pose_bwd_reg")) {
    c->use_pose_bwd_reg = value;
    return 0;
  }
  if (!strcmp(key, "compact_bwd")) {
    c->use_compact = value;
    return 0;
  }
  smplb_set_error("unknown debug key %s", key);
  return SMPLB_EINVAL;
}
// on = 1: per-kernel times (everything on one stream, so the events bracket one kernel each);
// on = 2: timeline trace -- the streams keep overlapping and smplb_profile_read returns one
// "name start_ms end_ms" line per launch, relative to a process-wide reference event, so the traces
// of several contexts of a device can be laid side by side (tools/timeline.py).
static cudaEvent_t g_trace_ref = nullptr;
extern "C" int smplb_profile_enable(smplb_ctx *c, int on) {
  RET_IF(!c, SMPLB_EINVAL, "null context");
  c->profile = on != 0;
  c->profile_serial = on == 1;   // per-kernel events are only meaningful without stream overlap
  c->profile_trace = on == 2;
  if (on == 2 && !g_trace_ref) {
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventCreate(&g_trace_ref));
    CUDA_TRY(cudaEventRecord(g_trace_ref, c->stream));
  }
  return 0;
}
extern "C" int smplb_profile_read(smplb_ctx *c, char *buf, size_t buflen) {
  CHECK_CTX(c);
  RET_IF(!buf || buflen == 0, SMPLB_EINVAL, "null buffer");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (c->stream2) CUDA_TRY(cudaStreamSynchronize(c->stream2));
  if (c->stream3) CUDA_TRY(cudaStreamSynchronize(c->stream3));
  if (c->stream_g) CUDA_TRY(cudaStreamSynchronize(c->stream_g));
  if (c->profile_trace && g_trace_ref) {
    std::string s;
    char line[256];
    for (auto &r : c->prof) {
      float t0 = 0.f, t1 = 0.f;
      cudaEventElapsedTime(&t0, g_trace_ref, r.e0);
      cudaEventElapsedTime(&t1, g_trace_ref, r.e1);
      snprintf(line, sizeof(line), "%s %.6f %.6f\n", r.name, t0, t1);
      s += line;
      c->event_pool.push_back(r.e0);
      c->event_pool.push_back(r.e1);
    }
    c->prof.clear();
    snprintf(buf, buflen, "%s", s.c_str());
    return 0;
  }
  std::map<std::string, std::pair<double, int>> acc;
  std::vector<std::string> order;
  for (auto &r : c->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (!acc.count(r.name)) order.push_back(r.name);
    acc[r.name].first += ms;
    acc[r.name].second += 1;
    c->event_pool.push_back(r.e0);
    c->event_pool.push_back(r.e1);
  }
  c->prof.clear();
  std::string s;
  char line[256];
  for (auto &n : order) {
    snprintf(line, sizeof(line), "%s %.6f %d\n", n.c_str(), acc[n].first, acc[n].second);
    s += line;
  }
  snprintf(buf, buflen, "%s", s.c_str());
  return 0;
}

// --------------------------------------------------------------------------------- SMPL fwd/bwd
// The 6890-vertex blend + skinning run on stream2 while the keypoint path continues on the
// main stream; whoever needs verts / v_posed on the main stream joins first.
// Error returns in the middle of a step leave work in flight on the side streams that still
// reads / writes staged buffers: drain them before anything is freed and forget the pending joins.
static void abort_side_streams(smplb_ctx *c) {
  for (cudaStream_t s : {c->stream2, c->stream3, c->stream_g, c->stream})
    if (s) cudaStreamSynchronize(s);
  c->verts_pending = false;
  c->gdx_pending = false;
  c->red_fork_recorded = false;
  c->cur = c->stream;
}
struct StepGuard {   // declared after the Stager, so it runs before the Stager frees its buffers
  smplb_ctx *c;
  bool ok = false;
  ~StepGuard() {
    if (!ok) abort_side_streams(c);
    c->cur = c->stream;   // error returns must not leave the context launching on a side stream
  }
};

static int join_verts(smplb_ctx *c) {
  if (c->verts_pending) {
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    c->verts_pending = false;
  }
  return 0;
}

static int smpl_forward_dev(smplb_ctx *c, int B, const float *beta, const float *theta, float *verts, float *joints,
                            float *Rs, float *Jtr, const float *cam, const float *kp_gt, float *kp_pred,
                            bool need_verts, bool want_vposed = false, float *step_d_cam = nullptr) {
  NvtxRange nvtx("smpl_main");   // batch_smpl.py:105
  TRY(ensure_ws(c, B));
  CUDA_TRY(cudaMemcpyAsync(c->ws_beta, beta, (size_t)B * c->NB * 4, cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->ws_theta, theta, (size_t)B * 72 * 4, cudaMemcpyDeviceToDevice, c->stream));
  bool tc = c->tc_ok && c->use_tc;
  bool stc = c->skin_tc_ok && c->use_skin_tc;
  bool fold = tc && c->fold_ok && c->use_fold;
  TRY(launch_pose_fwd(c, B, c->ws_beta, c->ws_theta, c->ws_Rs, c->ws_J, c->ws_A, Jtr ? Jtr : c->ws_Jtr,
                      tc ? nullptr : c->ws_x, tc ? c->ws_x16 : nullptr, stc ? c->ws_A16 : nullptr,
                      nullptr));   // (the 704-wide fold operand row is no longer used: the fold GEMM reads x16)
  if (Rs) CUDA_TRY(cudaMemcpyAsync(Rs, c->ws_Rs, (size_t)B * NJ * 9 * 4, cudaMemcpyDeviceToDevice, c->stream));
  // Keypoint path on the active vertices only (rows of joint_regressor with a non-zero): the
  // same two tensor-core kernels on ~9 % of the vertices give joints without reading verts back.
  bool compact = !fold && tc && stc && c->compact_ok && c->use_compact;
  bool full = need_verts || !(compact || fold);
  float *vout = verts;
  if (full && !vout) {
    TRY(ensure_buf(c, &c->ws_verts, (size_t)c->ws_batch * c->V3, false));
    vout = c->ws_verts;
  }
  // L2-resident hand-over: blend and skinning run chunk by chunk over the samples, the
  // v_posed of a chunk (chunk * 83 KB, 42 MB at 512) is written and read back inside the 126 MB
  // L2 and the same buffer is reused by the next chunk, so v_posed never makes the HBM round
  // trip.  The dense backward rebuilds v_posed when it needs it (saved_full = false).
  int chunk = c->l2_chunk;
  bool chunked = full && tc && stc && chunk > 0 && B > chunk;
  // One kernel for blend + skinning when nothing downstream is known to need v_posed (the dense
  // backward rebuilds it on demand, saved_full = false); the caller can ask for the two-kernel
  // path, which keeps v_posed, through want_vposed.
  bool fused = full && tc && stc && c->body_tc_ok && c->use_fused && !chunked && !want_vposed;
  // the fold GEMM needs TMEM and ~160 KB of shared memory, which the persistent blend / skinning
  // CTAs would deny it: issue it before forking so only the light per-body kernels overlap them
  bool overlap = full && (fold || compact) && c->use_overlap && !c->profile_serial;
  // (see stream_g)
  const bool prio = fold && c->use_prio && c->use_overlap && !c->profile_serial;
  if (fold && !prio) TRY(launch_fold_gemm_u(c, B, c->ws_x16));
  if (overlap || prio) CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
  if (prio) {
    // U = x G^T on the low-priority stream; the main stream resumes behind it
    CUDA_TRY(cudaStreamWaitEvent(c->stream_g, c->ev_fork, 0));
    c->cur = c->stream_g;
    int rc = launch_fold_gemm_u(c, B, c->ws_x16);
    c->cur = c->stream;
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c->ev_g_join, c->stream_g));
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_g_join, 0));
  }
  if (overlap) {
    CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    c->cur = c->stream2;
  }
  // A forward-only call: the keypoint path (one tcgen05 GEMM + a per-body kernel, ~25 us) runs on the main stream while
  // the vertex kernel runs here -- but only if that kernel leaves it SMs (the GEMM needs tensor memory and 160 KB of
  // shared memory, the vertex kernel takes all of both).  With 10 SM pairs left free the call takes 148 instead of
  // 185 us at B = 4096, 62 instead of 100 us at B = 1024, 39 instead of 78 us at B = 256 (tools/small_batch_pairs.py;
  // 66 pairs already lose it: 16 free SMs are not enough).  Not for the training step -- its longer keypoint chain does
  // not fit into 20 SMs beside the vertex kernel, and with several contexts in flight the SMs are never idle
  // (tools/pairs_vs_contexts.py) -- nor for batches whose vertex kernel dwarfs the keypoint path.
  c->pairs_auto = (overlap && !step_d_cam && B <= 8192 && c->num_sms >= 60) ? (c->num_sms - 20) / 2 : 0;
  if (fused) {
    TRY(launch_body_fwd_tc(c, B, c->ws_x16, c->ws_A16, vout));
  } else if (full && chunked) {
    for (int b0 = 0; b0 < B; b0 += chunk) {
      int nb = std::min(chunk, B - b0);
      TRY(launch_blend_fwd_tc(c, nb, (const char *)c->ws_x16 + (size_t)b0 * 512, c->ws_vposed, false));
      TRY(launch_skin_fwd_tc(c, nb, (const char *)c->ws_A16 + (size_t)b0 * 12 * 128, c->ws_vposed,
                             vout + (size_t)b0 * c->V3, false));
    }
  } else if (full) {
    if (tc) TRY(launch_blend_fwd_tc(c, B, c->ws_x16, c->ws_vposed, false));
    else TRY(launch_blend_fwd(c, B, c->ws_x, c->ws_vposed));
    if (stc) TRY(launch_skin_fwd_tc(c, B, c->ws_A16, c->ws_vposed, vout, false));
    else TRY(launch_skin_fwd(c, B, c->ws_A, c->ws_vposed, vout));
  }
  c->saved_full = full && !chunked && !fused;
  if (overlap) {
    cudaError_t e1 = cudaEventRecord(c->ev_join, c->stream2);
    c->cur = c->stream;
    c->verts_pending = true;
    if (e1 != cudaSuccess) {
      smplb_set_error("cudaEventRecord(ev_join) failed: %s", cudaGetErrorString(e1));
      return SMPLB_ECUDA;
    }
  }
  c->saved_verts = vout;
  float *jout = joints ? joints : c->ws_joints;
  c->saved_fold = fold;
  c->saved_fold_step = false;
  if (fold && step_d_cam && cam && kp_gt && c->fold_warp_kernels && c->use_fold_step && c->K * NJ * 3 <= 1536) {
    // smplb_step with the keypoint loss only: forward and backward of the folded path in one
    // kernel + the dx GEMM; gradients for a unit loss scale (k_pose_bwd applies w_kp / num_present)
    TRY(launch_fold_step(c, B, c->ws_A, cam, kp_gt, jout, kp_pred, c->ws_part, c->ws_cnt, step_d_cam, c->ws_dA, c->ws_dx, 2));
    c->saved_fold_step = true;
  } else if (fold) {
    // joints from x and A alone (k_fold.cu): one small GEMM + a per-body contraction
    TRY(launch_fold_fwd(c, B, c->ws_x16b, c->ws_A, cam, kp_gt, jout, kp_pred, kp_gt ? c->ws_dkp : nullptr,
                        kp_gt ? c->ws_part : nullptr, kp_gt ? c->ws_cnt : nullptr));
  } else if (compact) {
    TRY(ensure_buf(c, &c->ws_vposed_act, (size_t)c->ws_batch * c->pitch_act, false));
    TRY(ensure_buf(c, &c->ws_verts_act, (size_t)c->ws_batch * c->n_act * 3, false));
    TRY(launch_blend_fwd_tc(c, B, c->ws_x16, c->ws_vposed_act, true));
    TRY(launch_skin_fwd_tc(c, B, c->ws_A16, c->ws_vposed_act, c->ws_verts_act, true));
    TRY(launch_joints(c, B, c->ws_verts_act, cam, kp_gt, jout, kp_pred, kp_gt ? c->ws_dkp : nullptr,
                      kp_gt ? c->ws_part : nullptr, kp_gt ? c->ws_cnt : nullptr, true));
  } else {
    TRY(launch_joints(c, B, vout, cam, kp_gt, jout, kp_pred, kp_gt ? c->ws_dkp : nullptr, kp_gt ? c->ws_part : nullptr,
                      kp_gt ? c->ws_cnt : nullptr, false));
  }
  c->saved_compact = compact;
  c->saved_B = B;
  return 0;
}

static int smpl_backward_dev(smplb_ctx *c, int B, const float *d_verts, const float *d_joints, const float *d_Rs,
                             float *d_beta, float *d_theta) {
  RET_IF(c->saved_B != B, SMPLB_ESTATE, "smplb_smpl_backward(B=%d) without a matching forward (saved B=%d)", B,
         c->saved_B);
  if (d_verts != nullptr || !c->saved_fold) TRY(join_verts(c));   // the per-vertex backward reads v_posed
  if (d_verts == nullptr && d_joints != nullptr && c->saved_fold) {
    // gradient arrives through the keypoints only: folded backward, no per-vertex work
    int rows = cdiv(B, 128) * 128;
    TRY(launch_fold_bwd(c, B, c->ws_A, d_joints, nullptr, nullptr, nullptr, 1.0f, nullptr, nullptr, c->ws_dA, c->ws_dx, 2));
    TRY(launch_pose_bwd(c, B, c->ws_theta, c->ws_Rs, c->ws_J, c->ws_A, c->ws_dA, 1, c->ws_dx, 2, rows, c->ws_rowscale, d_Rs,
                        d_beta, d_theta));
    return 0;
  }
  // No upstream gradient on verts: only vertices the keypoint regressor touches carry a
  // gradient, so walk just those (exact: the skipped terms are zeros).
  bool compact = (d_verts == nullptr) && c->use_compact && c->n_act < c->V;
  float *dp = nullptr;
  const bool bwd_tc = !compact && c->tc_ok && c->use_blend_bwd_tc;
  if (compact) {
    TRY(ensure_buf(c, &c->ws_dp_act, (size_t)c->ws_batch * c->pitch_act, true));
    dp = c->ws_dp_act;
  } else if (!bwd_tc) {
    TRY(ensure_buf(c, &c->ws_dp, (size_t)c->ws_batch * c->pitch, true));
    dp = c->ws_dp;
  }
  if (bwd_tc) {
    // k_skin_bwd writes the bf16 operand rows of the tcgen05 blend-transpose GEMM instead of fp32 dp
    if (!c->ws_dp16) {
      CUDA_TRY(cudaMalloc(&c->ws_dp16, (size_t)c->ws_batch * 3 * c->pitch * 2));
      CUDA_TRY(cudaMemsetAsync(c->ws_dp16, 0, (size_t)c->ws_batch * 3 * c->pitch * 2, c->cur));   // padding columns stay 0
    }
  }
  if ((!compact || !c->saved_compact) && !c->saved_full) {
    // the forward skipped the 6890-vertex tensors (no verts requested): rebuild v_posed now
    if (c->tc_ok && c->use_tc) TRY(launch_blend_fwd_tc(c, B, c->ws_x16, c->ws_vposed, false));
    else RET_IF(true, SMPLB_ESTATE, "v_posed was not saved by the forward");
    c->saved_full = true;
  }
  int dA_parts = skin_bwd_splits(B);
  if (compact && c->saved_compact)
    TRY(launch_skin_bwd(c, B, c->ws_A, c->ws_vposed_act, nullptr, d_joints, dp, c->ws_dA, 2));
  else if (!compact && c->skin_bwd_tc_ok && c->use_skin_bwd_tc && c->skin_tc_ok && c->use_skin_tc)
    // the dense walk with both contractions on the tensor cores (A16 is the forward's operand, still in the workspace)
    TRY(launch_skin_bwd_tc(c, B, c->ws_A16, c->ws_vposed, d_verts, d_joints, dp, bwd_tc ? c->ws_dp16 : nullptr, c->ws_dA, &dA_parts));
  else
    TRY(launch_skin_bwd(c, B, c->ws_A, c->ws_vposed, d_verts, d_joints, dp, c->ws_dA, compact ? 1 : 0, bwd_tc ? c->ws_dp16 : nullptr));
  int ks = compact ? 4 : c->ksplit;   // the compact contraction is 11x shorter: fewer split-K partials
  int dx_rows = B;
  if (bwd_tc) TRY(launch_blend_bwd_tc(c, B, c->ws_dp16, c->ws_dx, &ks, &dx_rows));
  else TRY(launch_blend_bwd(c, B, dp, c->ws_dx, compact, ks));
  TRY(launch_pose_bwd(c, B, c->ws_theta, c->ws_Rs, c->ws_J, c->ws_A, c->ws_dA, dA_parts, c->ws_dx, ks, dx_rows, nullptr, d_Rs,
                      d_beta, d_theta));
  return 0;
}

extern "C" int smplb_smpl_forward(smplb_ctx *c, int B, const float *beta, const float *theta, float *verts,
                                  float *joints, float *Rs, float *J_transformed, int flags, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(flags & ~SMPLB_STEP_KEEP_VERTS, SMPLB_EINVAL, "unknown bits in flags");
  RET_IF(B < 1 || !beta || !theta || !joints, SMPLB_EINVAL, "B >= 1 and non-null beta, theta, joints required");
  Stager st(c, mem);
  const float *dbeta = st.in(beta, (size_t)B * c->NB), *dtheta = st.in(theta, (size_t)B * 72);
  float *dverts = st.out(verts, (size_t)B * c->V3), *djoints = st.out(joints, (size_t)B * c->K * 3);
  float *dRs = st.out(Rs, (size_t)B * NJ * 9), *dJtr = st.out(J_transformed, (size_t)B * NJ * 3);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  StepGuard guard{c};
  TRY(smpl_forward_dev(c, B, dbeta, dtheta, dverts, djoints, dRs, dJtr, nullptr, nullptr, nullptr,
                       verts != nullptr || (flags & SMPLB_STEP_KEEP_VERTS)));
  TRY(join_verts(c));
  int rc = st.finish();
  guard.ok = rc == 0;
  return rc;
}

extern "C" int smplb_smpl_backward(smplb_ctx *c, int B, const float *d_verts, const float *d_joints,
                                   const float *d_Rs, float *d_beta, float *d_theta, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(B < 1 || !d_beta || !d_theta, SMPLB_EINVAL, "B >= 1 and non-null d_beta, d_theta required");
  Stager st(c, mem);
  const float *dv = st.in(d_verts, (size_t)B * c->V3), *dj = st.in(d_joints, (size_t)B * c->K * 3);
  const float *dr = st.in(d_Rs, (size_t)B * NJ * 9);
  float *db = st.out(d_beta, (size_t)B * c->NB), *dt = st.out(d_theta, (size_t)B * 72);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(smpl_backward_dev(c, B, dv, dj, dr, db, dt));
  return st.finish();
}

extern "C" int smplb_last_verts(smplb_ctx *c, const float **dptr) {
  RET_IF(!c || !dptr, SMPLB_EINVAL, "null argument");
  *dptr = c->saved_full || c->saved_verts ? c->saved_verts : nullptr;
  return 0;
}

extern "C" int smplb_rodrigues(smplb_ctx *c, int N, const float *theta, float *R, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("batch_rodrigues");   // the reference's tf.name_scope of this function
  RET_IF(N < 1 || !theta || !R, SMPLB_EINVAL, "N >= 1 and non-null pointers required");
  Stager st(c, mem);
  const float *dt = st.in(theta, (size_t)N * 3);
  float *dR = st.out(R, (size_t)N * 9);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_rodrigues(c, N, dt, dR));
  return st.finish();
}

extern "C" int smplb_global_rigid(smplb_ctx *c, int B, const float *Rs, const float *Js, float *new_J, float *A,
                                  int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("batch_forward_kinematics");   // the reference's tf.name_scope of this function
  RET_IF(B < 1 || !Rs || !Js || !new_J || !A, SMPLB_EINVAL, "B >= 1 and non-null pointers required");
  Stager st(c, mem);
  const float *dR = st.in(Rs, (size_t)B * NJ * 9), *dJ = st.in(Js, (size_t)B * NJ * 3);
  float *dn = st.out(new_J, (size_t)B * NJ * 3), *dA = st.out(A, (size_t)B * NJ * 16);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_global_rigid(c, B, dR, dJ, dn, dA));
  return st.finish();
}

// --------------------------------------------------------------------------------- projection
extern "C" int smplb_orth_proj(smplb_ctx *c, int B, int N, const float *X, const float *cam, float *out, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("batch_orth_proj_idrot");   // the reference's tf.name_scope of this function
  RET_IF(B < 1 || N < 1 || !X || !cam || !out, SMPLB_EINVAL, "B, N >= 1 and non-null pointers required");
  Stager st(c, mem);
  const float *dX = st.in(X, (size_t)B * N * 3), *dc = st.in(cam, (size_t)B * 3);
  float *dout = st.out(out, (size_t)B * N * 2);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_proj(c, B, N, dX, dc, 0, 0.f, 0.f, dout));
  return st.finish();
}

extern "C" int smplb_reproject_vertices(smplb_ctx *c, int B, int N, const float *verts, const float *cam, float im_w,
                                        float im_h, float *out, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("mesh_reproject");   // the reference's tf.name_scope of this function
  RET_IF(B < 1 || N < 1 || !verts || !cam || !out, SMPLB_EINVAL, "B, N >= 1 and non-null pointers required");
  Stager st(c, mem);
  const float *dX = st.in(verts, (size_t)B * N * 3), *dc = st.in(cam, (size_t)B * 3);
  float *dout = st.out(out, (size_t)B * N * 2);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_proj(c, B, N, dX, dc, 1, im_w, im_h, dout));
  return st.finish();
}

extern "C" int smplb_proj_backward(smplb_ctx *c, int B, int N, const float *X, const float *cam, const float *d_out,
                                   int pixel, float im_w, float im_h, float *d_X, float *d_cam, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(B < 1 || N < 1 || !X || !cam || !d_out, SMPLB_EINVAL, "B, N >= 1 and non-null X, cam, d_out required");
  Stager st(c, mem);
  const float *dX = st.in(X, (size_t)B * N * 3), *dc = st.in(cam, (size_t)B * 3), *dd = st.in(d_out, (size_t)B * N * 2);
  float *oX = st.out(d_X, (size_t)B * N * 3), *oc = st.out(d_cam, (size_t)B * 3);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_proj_bwd(c, B, N, dX, dc, dd, pixel, im_w, im_h, 1.0f, nullptr, 0, oX, oc));
  return st.finish();
}

// -------------------------------------------------------------------------------------- losses
extern "C" int smplb_kp_loss(smplb_ctx *c, int B, int K, const float *kp_gt, const float *kp_pred, float *abs_sum,
                             int64_t *num_present, float *d_kp_pred, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("kp_reprojection_loss");   // the reference's tf.name_scope of this function
  RET_IF(B < 1 || K < 1 || !kp_gt || !kp_pred || !abs_sum || !num_present, SMPLB_EINVAL,
         "B, K >= 1 and non-null kp_gt, kp_pred, abs_sum, num_present required");
  TRY(ensure_ws(c, B));
  Stager st(c, mem);
  const float *dg = st.in(kp_gt, (size_t)B * K * 3), *dp = st.in(kp_pred, (size_t)B * K * 2);
  float *ds = st.out(abs_sum, 1);
  long long *dn = (long long *)st.out(num_present, 1);
  float *dd = st.out(d_kp_pred, (size_t)B * K * 2);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_kp_loss(c, B, K, dg, dp, dd, c->ws_part, c->ws_cnt));
  TRY(launch_reduce_kp(c, B, c->ws_part, c->ws_cnt, ds, dn, nullptr));
  return st.finish();
}

static int ensure_mesh_ws(smplb_ctx *c, int B, int V) {
  size_t n = (size_t)std::max(B, c->ws_batch) * std::max(V, c->V) * 2;
  if (n > c->ws_mesh_cap) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (void **p : {(void **)&c->ws_silpred, (void **)&c->ws_dsil, (void **)&c->ws_silcnt}) {
      if (*p) CUDA_TRY(cudaFree(*p));
      *p = nullptr;
      CUDA_TRY(cudaMalloc(p, n * 4));
    }
    c->ws_mesh_cap = n;
  }
  size_t np = (size_t)B * (32 + cdiv(V, 256) + 1);
  if (np > c->ws_mesh_part_cap) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->ws_mesh_part) CUDA_TRY(cudaFree(c->ws_mesh_part));
    c->ws_mesh_part = nullptr;
    CUDA_TRY(cudaMalloc((void **)&c->ws_mesh_part, np * 4));
    c->ws_mesh_part_cap = np;
  }
  return 0;
}

extern "C" int smplb_mesh_reproj_loss(smplb_ctx *c, int B, int V, const float *points_xy, const int32_t *offsets, int P,
                                      const float *sil_pred, float *loss, float *d_sil_pred, int32_t *ind_ab,
                                      int32_t *ind_ba, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("mesh_reprojection_loss");   // the reference's tf.name_scope of this function
  RET_IF(B < 1 || V < 1 || P < 0 || !offsets || !sil_pred || !loss || (P > 0 && !points_xy), SMPLB_EINVAL,
         "B, V >= 1, P >= 0 and non-null offsets, sil_pred, loss required");
  TRY(ensure_mesh_ws(c, B, V));
  Stager st(c, mem);
  const float *dp = st.in(points_xy, (size_t)P * 2);
  const int32_t *dof = st.in(offsets, (size_t)B + 1);
  const float *ds = st.in(sil_pred, (size_t)B * V * 2);
  float *dl = st.out(loss, 1), *dg = st.out(d_sil_pred, (size_t)B * V * 2);
  int32_t *dia = st.out(ind_ab, (size_t)P), *dib = st.out(ind_ba, (size_t)B * V);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  if (!dp) dp = c->ws_scal;  // P == 0: never dereferenced
  TRY(launch_mesh_loss(c, B, V, dp, dof, P, ds, dl, dg, c->ws_silcnt, c->ws_mesh_part, dia, dib));
  return st.finish();
}

extern "C" int smplb_skew(smplb_ctx *c, int N, const float *vec, float *out, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("batch_skew");   // the reference's tf.name_scope of this function
  RET_IF(N < 1 || !vec || !out, SMPLB_EINVAL, "N >= 1 and non-null pointers required");
  Stager st(c, mem);
  const float *dv = st.in(vec, (size_t)N * 3);
  float *d_o = st.out(out, (size_t)N * 9);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_skew(c, N, dv, d_o));
  return st.finish();
}

extern "C" int smplb_lrotmin(smplb_ctx *c, int B, const float *theta, float *out, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("batch_lrotmin");   // the reference's tf.name_scope of this function
  RET_IF(B < 1 || !theta || !out, SMPLB_EINVAL, "B >= 1 and non-null pointers required");
  Stager st(c, mem);
  const float *dt = st.in(theta, (size_t)B * 72);
  float *d_o = st.out(out, (size_t)B * NPF);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_lrotmin(c, B, dt, d_o));
  return st.finish();
}

static int ensure_gp_ws(smplb_ctx *c, int M) {
  size_t need = ((size_t)cdiv(M, 64) + 2) * SMPLB_GP_FLOATS;
  if (need <= c->ws_gp_cap) return 0;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (c->ws_gp) CUDA_TRY(cudaFree(c->ws_gp));
  c->ws_gp = nullptr;
  CUDA_TRY(cudaMalloc((void **)&c->ws_gp, need * 4));
  c->ws_gp_cap = need;
  return 0;
}

extern "C" int smplb_gradient_penalty(smplb_ctx *c, int M, const float *g0, const float *g1, const float *g2,
                                      const float *g3, float *penalty, float *col_sums, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("compute_gradient_penalty");   // the reference's tf.name_scope of this function
  RET_IF(M < 1 || !g0 || !g1 || !g2 || !g3 || !penalty, SMPLB_EINVAL, "M >= 1 and non-null g0..g3, penalty required");
  TRY(ensure_gp_ws(c, M));
  Stager st(c, mem);
  const float *d0 = st.in(g0, (size_t)M * 169), *d1 = st.in(g1, (size_t)M * 42), *d2 = st.in(g2, (size_t)M * 10),
              *d3 = st.in(g3, (size_t)M * 207);
  float *dp = st.out(penalty, 1), *dc = st.out(col_sums, SMPLB_GP_FLOATS);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  float *sums = dc ? dc : c->ws_gp + (size_t)cdiv(M, 64) * SMPLB_GP_FLOATS;
  TRY(launch_gp_colsum(c, M, d0, d1, d2, d3, sums));
  TRY(launch_gp_final(c, (long long)M, sums, dp));
  return st.finish();
}

extern "C" int smplb_gradient_penalty_from_sums(smplb_ctx *c, int64_t M_total, const float *col_sums, float *penalty,
                                                int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(M_total < 1 || !col_sums || !penalty, SMPLB_EINVAL, "M_total >= 1 and non-null pointers required");
  Stager st(c, mem);
  const float *dc = st.in(col_sums, SMPLB_GP_FLOATS);
  float *dp = st.out(penalty, 1);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_gp_final(c, (long long)M_total, dc, dp));
  return st.finish();
}

extern "C" int smplb_gradient_penalty_backward(smplb_ctx *c, int M, int64_t M_total, const float *col_sums, float *d_g0,
                                               float *d_g1, float *d_g2, float *d_g3, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(M < 1 || M_total < 1 || !col_sums, SMPLB_EINVAL, "M, M_total >= 1 and non-null col_sums required");
  Stager st(c, mem);
  const float *dc = st.in(col_sums, SMPLB_GP_FLOATS);
  float *o0 = st.out(d_g0, (size_t)M * 169), *o1 = st.out(d_g1, (size_t)M * 42), *o2 = st.out(d_g2, (size_t)M * 10),
        *o3 = st.out(d_g3, (size_t)M * 207);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_gp_bwd(c, M, (long long)M_total, dc, o0, o1, o2, o3));
  return st.finish();
}

// ------------------------------------------------------------------- SURVEY section 8f rows
extern "C" int smplb_silhouette_csr(smplb_ctx *c, int B, int H, int W, const float *seg, float *points_xy, int cap,
                                    int32_t *offsets, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(B < 1 || H < 1 || W < 1 || cap < 0 || !seg || !offsets || (cap > 0 && !points_xy), SMPLB_EINVAL,
         "B, H, W >= 1 and non-null seg, offsets (and points_xy when cap > 0) required");
  TRY(ensure_ws(c, B));
  Stager st(c, mem);
  const float *ds = st.in(seg, (size_t)B * H * W);
  float *dp = st.out(points_xy, (size_t)cap * 2);
  int32_t *dof = st.out(offsets, (size_t)B + 1);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  if (!dp) dp = c->ws_scal;
  TRY(launch_silhouette_csr(c, B, H, W, ds, dp, cap, dof, c->ws_cnt));
  return st.finish();
}

extern "C" int smplb_kcs(smplb_ctx *c, int N, int K, const float *joints, const float *Cm, float *kcs, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  NvtxRange nvtx("get_kcs");   // the reference's tf.name_scope of this function
  RET_IF(N < 1 || K < 14 || !joints || !Cm || !kcs, SMPLB_EINVAL, "N >= 1, K >= 14 and non-null pointers required");
  Stager st(c, mem);
  const float *dj = st.in(joints, (size_t)N * K * 3), *dc = st.in(Cm, (size_t)14 * 13);
  float *dk = st.out(kcs, (size_t)N * 169);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_kcs(c, N, K, dj, dc, dk));
  return st.finish();
}

extern "C" int smplb_kcs_backward(smplb_ctx *c, int N, int K, const float *joints, const float *Cm, const float *d_kcs,
                                  float *d_joints, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(N < 1 || K < 14 || !joints || !Cm || !d_kcs || !d_joints, SMPLB_EINVAL,
         "N >= 1, K >= 14 and non-null pointers required");
  Stager st(c, mem);
  const float *dj = st.in(joints, (size_t)N * K * 3), *dc = st.in(Cm, (size_t)14 * 13), *dk = st.in(d_kcs, (size_t)N * 169);
  float *o = st.out(d_joints, (size_t)N * K * 3);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_kcs_bwd(c, N, K, dj, dc, dk, o));
  return st.finish();
}

extern "C" int smplb_interpolate(smplb_ctx *c, int N, int row, const float *fake, const float *real, const float *alpha,
                                 float *out, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(N < 1 || row < 1 || !fake || !real || !alpha || !out, SMPLB_EINVAL, "N, row >= 1 and non-null pointers required");
  Stager st(c, mem);
  const float *df = st.in(fake, (size_t)N * row), *dr = st.in(real, (size_t)N * row), *da = st.in(alpha, (size_t)N);
  float *o = st.out(out, (size_t)N * row);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_interp(c, (size_t)N * row, row, df, dr, da, o));
  return st.finish();
}

// trainer.py:548-557 in one launch: element-wise interpolation of the critic's three inputs + get_kcs of the result
extern "C" int smplb_critic_inputs(smplb_ctx *c, int N, int K, const float *fake_joints, const float *real_joints,
                                   const float *alpha_joints, const float *fake_shapes, const float *real_shapes,
                                   const float *alpha_shapes, const float *fake_Rs, const float *real_Rs,
                                   const float *alpha_Rs, const float *Cm, float *joints, float *kcs, float *shapes,
                                   float *Rs, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(N < 1 || K < 14 || !fake_joints || !real_joints || !alpha_joints || !fake_shapes || !real_shapes || !alpha_shapes ||
             !fake_Rs || !real_Rs || !alpha_Rs || !Cm || !joints || !kcs || !shapes || !Rs,
         SMPLB_EINVAL, "N >= 1, K >= 14 and non-null pointers required");
  NvtxRange nvtx("critic_inputs");
  Stager st(c, mem);
  const size_t nj = (size_t)N * K * 3, ns = (size_t)N * 10, nR = (size_t)N * 207;
  const float *fj = st.in(fake_joints, nj), *rj = st.in(real_joints, nj), *aj = st.in(alpha_joints, nj);
  const float *fs = st.in(fake_shapes, ns), *rs = st.in(real_shapes, ns), *as_ = st.in(alpha_shapes, ns);
  const float *fR = st.in(fake_Rs, nR), *rR = st.in(real_Rs, nR), *aR = st.in(alpha_Rs, nR);
  const float *dc = st.in(Cm, (size_t)14 * 13);
  float *oj = st.out(joints, nj), *ok = st.out(kcs, (size_t)N * 169), *os = st.out(shapes, ns), *oR = st.out(Rs, nR);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  TRY(launch_critic_inputs(c, N, K, fj, rj, aj, fs, rs, as_, fR, rR, aR, dc, oj, ok, os, oR));
  return st.finish();
}

// trainer.py:566-572 + ops.py:153-172 in one launch, given the critic's partial derivatives
extern "C" int smplb_critic_gradient_penalty(smplb_ctx *c, int M, int K, int64_t M_total, const float *joints, const float *Cm,
                                             const float *g_kcs, const float *g_joints, const float *g_shapes,
                                             const float *g_Rs, float *penalty, float *col_sums, float *g_joints_total,
                                             int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(M < 1 || K < 14 || M_total < M || !joints || !Cm || !g_kcs || !g_joints || !g_shapes || !g_Rs || !penalty, SMPLB_EINVAL,
         "M >= 1, K >= 14, M_total >= M and non-null joints, C, gradients, penalty required");
  NvtxRange nvtx("compute_gradient_penalty");
  TRY(ensure_gp_ws(c, M));
  if (!c->ws_ticket) {
    CUDA_TRY(cudaMalloc((void **)&c->ws_ticket, 4));
    CUDA_TRY(cudaMemsetAsync(c->ws_ticket, 0, 4, c->stream));
  }
  Stager st(c, mem);
  const float *dj = st.in(joints, (size_t)M * K * 3), *dc = st.in(Cm, (size_t)14 * 13), *gk = st.in(g_kcs, (size_t)M * 169);
  const float *gj = st.in(g_joints, (size_t)M * 42), *gs = st.in(g_shapes, (size_t)M * 10), *gR = st.in(g_Rs, (size_t)M * 207);
  float *dp = st.out(penalty, 1), *dcs = st.out(col_sums, SMPLB_GP_FLOATS), *dgt = st.out(g_joints_total, (size_t)M * 42);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  float *sums = dcs ? dcs : c->ws_gp + (size_t)cdiv(M, 64) * SMPLB_GP_FLOATS;
  TRY(launch_critic_gp(c, M, K, (long long)M_total, dj, dc, gk, gj, gs, gR, dgt, c->ws_gp, c->ws_ticket, sums, dp));
  return st.finish();
}

// ---------------------------------------------------------------------------------- fused step
static int allreduce_on(smplb_ctx *c, void *dev_buf, size_t count, int nccl_dtype, cudaStream_t stream);

// seg != NULL: the silhouettes arrive as the dense mask seg [B,H,W] (src/trainer.py:443: where(seg > 0)) and are
// compacted on the device into the context's workspace; else points_xy / offsets / P as documented in smplb.h.
static int step_core(smplb_ctx *c, int B, const float *beta, const float *theta, const float *cam, const float *kp_gt,
                     const float *points_xy, const int32_t *offsets, int P, const float *seg, int H, int W, float w_kp,
                     float w_mesh, float img_size, int64_t kp_count_override, float *verts, float *joints, float *Rs,
                     float *kp_pred, float *loss_parts, float *d_beta, float *d_theta, float *d_cam, int flags, int mem) {
  CHECK_CTX(c);
  CHECK_MEM(mem);
  RET_IF(B < 1 || !beta || !theta || !cam || !kp_gt || !loss_parts, SMPLB_EINVAL,
         "B >= 1 and non-null beta, theta, cam, kp_gt, loss_parts required");
  RET_IF(flags & ~SMPLB_STEP_KEEP_VERTS, SMPLB_EINVAL, "unknown bits in flags");
  RET_IF(seg && (H < 1 || W < 1), SMPLB_EINVAL, "H, W >= 1 required with a dense mask");
  bool have_mesh = offsets != nullptr || seg != nullptr;
  bool bwd = d_beta || d_theta || d_cam;
  RET_IF(bwd && !(d_beta && d_theta && d_cam), SMPLB_EINVAL, "d_beta, d_theta, d_cam must be all set or all NULL");
  RET_IF(!seg && have_mesh && P > 0 && !points_xy, SMPLB_EINVAL, "points_xy is NULL but P > 0");
  TRY(ensure_ws(c, B));
  if (have_mesh) TRY(ensure_mesh_ws(c, B, c->V));
  int K = c->K;
  NvtxRange nvtx_step("smplb_step");
  Stager st(c, mem);
  StepGuard guard{c};
  const float *dbeta = st.in(beta, (size_t)B * c->NB), *dtheta = st.in(theta, (size_t)B * 72);
  const float *dcam = st.in(cam, (size_t)B * 3), *dkpgt = st.in(kp_gt, (size_t)B * K * 3);
  const float *dpts = (have_mesh && !seg) ? st.in(points_xy, (size_t)P * 2) : nullptr;
  const int32_t *doff = (have_mesh && !seg) ? st.in(offsets, (size_t)B + 1) : nullptr;
  const float *dseg = seg ? st.in(seg, (size_t)B * H * W) : nullptr;
  float *overts = st.out(verts, (size_t)B * c->V3), *ojoints = st.out(joints, (size_t)B * K * 3);
  float *oRs = st.out(Rs, (size_t)B * NJ * 9), *okp = st.out(kp_pred, (size_t)B * K * 2);
  float *oloss = st.out(loss_parts, 4);
  float *odb = st.out(d_beta, (size_t)B * c->NB), *odt = st.out(d_theta, (size_t)B * 72), *odc = st.out(d_cam, (size_t)B * 3);
  RET_IF(st.failed, SMPLB_ECUDA, "device staging allocation failed");
  if (seg) {
    // where(seg > 0) -> CSR point lists, on the device (k_extra.cu): counts and offsets first, one 4-byte read-back
    // of the total so the point buffer is sized exactly, then the ordered fill
    if ((size_t)B + 1 > c->ws_segoff_cap) {
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      if (c->ws_segoff) CUDA_TRY(cudaFree(c->ws_segoff));
      c->ws_segoff = nullptr;
      CUDA_TRY(cudaMalloc((void **)&c->ws_segoff, ((size_t)B + 1) * 4));
      c->ws_segoff_cap = (size_t)B + 1;
    }
    TRY(launch_silhouette_csr(c, B, H, W, dseg, c->ws_scal /* unused: cap 0 */, 0, c->ws_segoff, c->ws_cnt));
    int total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, c->ws_segoff + B, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    P = total;
    if ((size_t)P > c->ws_segpts_cap) {
      if (c->ws_segpts) CUDA_TRY(cudaFree(c->ws_segpts));
      c->ws_segpts = nullptr;
      CUDA_TRY(cudaMalloc((void **)&c->ws_segpts, std::max<size_t>((size_t)P, 1) * 8));
      c->ws_segpts_cap = (size_t)P;
    }
    if (P > 0) TRY(launch_silhouette_fill(c, B, H, W, dseg, c->ws_segpts, P, c->ws_segoff));
    dpts = c->ws_segpts;
    doff = c->ws_segoff;
  }
  if (have_mesh && !dpts) dpts = c->ws_scal;

  // ---- batch shards: the visibility count is exchanged first (it depends on kp_gt alone, SURVEY
  //      section 8e), on stream3, so the backward never waits for a peer.  ws_cnt64[0] = the
  //      denominator of this step's keypoint gradients; ws_cnt64[1] = the local count.
  const bool p2p = c->x_attached && !(c->comm_backend == 1 && c->nccl_comm);
  const bool nccl = !p2p && c->nccl_comm != nullptr;
  const bool comm = p2p || nccl;
  long long *den = c->ws_cnt64;
  bool cnt_aside = false;
  if (comm) {
    c->x_epoch++;
    cnt_aside = c->use_overlap && !c->profile_serial;
    if (cnt_aside) {
      CUDA_TRY(cudaEventRecord(c->ev_step0, c->stream));
      CUDA_TRY(cudaStreamWaitEvent(c->stream3, c->ev_step0, 0));
      c->cur = c->stream3;
    }
    TRY(launch_count_exchange(c, B, dkpgt, (long long)kp_count_override, p2p ? 1 : 2, den));
    if (nccl && kp_count_override <= 0) TRY(allreduce_on(c, den, 1, /*ncclInt64*/ 4, c->cur));
    if (cnt_aside) {
      CUDA_TRY(cudaEventRecord(c->ev_cnt, c->stream3));
      c->cur = c->stream;
    }
  }

  const bool keep = (flags & SMPLB_STEP_KEEP_VERTS) != 0;
  float *jbuf = ojoints ? ojoints : c->ws_joints;
  TRY(smpl_forward_dev(c, B, dbeta, dtheta, overts, jbuf, oRs, nullptr, dcam, dkpgt, okp ? okp : c->ws_kp,
                       overts != nullptr || have_mesh || keep, /*want_vposed=*/have_mesh && bwd,
                       /*step_d_cam=*/(bwd && !have_mesh) ? odc : nullptr));
  const float *vbuf = c->saved_verts;
  // Keypoint step: the loss reduction (and the exchange of the numerators) only needs what
  // k_fold_step_w wrote, so it runs on stream3 next to the dx GEMM that was queued behind that
  // kernel; the main stream picks the result up again below.
  const bool red_aside = c->red_fork_recorded && c->saved_fold_step && !have_mesh && !c->profile_serial;
  c->red_fork_recorded = false;
  if (red_aside) {
    CUDA_TRY(cudaStreamWaitEvent(c->stream3, c->ev_red_fork, 0));
    c->cur = c->stream3;
  }
  if (!have_mesh && !comm) {
    TRY(launch_reduce_finalize(c, B, w_kp, w_mesh, (long long)kp_count_override, oloss));
  } else if (!comm) {
    TRY(launch_reduce_kp(c, B, c->ws_part, c->ws_cnt, c->ws_scal + 0, c->ws_cnt64, c->ws_scal + 1));
    TRY(join_verts(c));
    TRY(launch_proj(c, B, c->V, vbuf, dcam, 1, img_size, img_size, c->ws_silpred));
    TRY(launch_mesh_loss(c, B, c->V, dpts, doff, P, c->ws_silpred, c->ws_scal + 2, bwd ? c->ws_dsil : nullptr,
                         c->ws_silcnt, c->ws_mesh_part, nullptr, nullptr, /*finish_grad=*/false));
    TRY(launch_finalize_loss(c, w_kp, w_mesh, (long long)kp_count_override, 1, oloss));
  } else {
    // the path's one exchange: {kp numerator, mesh sum} summed over the batch shards, everything on
    // c->cur (stream3 for the keypoint step) -- reduce, exchange and finalize share ONE stream
    if (have_mesh) {
      TRY(launch_reduce_kp(c, B, c->ws_part, c->ws_cnt, c->ws_scal + 0, c->ws_cnt64 + 1, c->ws_scal + 1));
      TRY(join_verts(c));
      TRY(launch_proj(c, B, c->V, vbuf, dcam, 1, img_size, img_size, c->ws_silpred));
      TRY(launch_mesh_loss(c, B, c->V, dpts, doff, P, c->ws_silpred, c->ws_scal + 2, bwd ? c->ws_dsil : nullptr,
                           c->ws_silcnt, c->ws_mesh_part, nullptr, nullptr, /*finish_grad=*/false));
    }
    if (cnt_aside && c->cur != c->stream3) {
      CUDA_TRY(cudaStreamWaitEvent(c->cur, c->ev_cnt, 0));   // den (written on stream3) before its first reader here
      cnt_aside = false;
    }
    if (p2p) {
      TRY(launch_reduce_exchange_finalize(c, B, have_mesh ? nullptr : c->ws_part, w_kp, w_mesh, have_mesh ? 1 : 0, den,
                                          oloss));
    } else {
      if (!have_mesh) {
        TRY(launch_reduce_kp(c, B, c->ws_part, c->ws_cnt, c->ws_scal + 0, c->ws_cnt64 + 1, c->ws_scal + 1));
        CUDA_TRY(cudaMemsetAsync(c->ws_scal + 2, 0, 4, c->cur));
      }
      TRY(allreduce_on(c, c->ws_scal, 3, /*ncclFloat*/ 7, c->cur));
      TRY(launch_finalize_den(c, w_kp, w_mesh, have_mesh ? 1 : 0, den, oloss));
    }
  }
  if (red_aside) {
    CUDA_TRY(cudaEventRecord(c->ev_red_join, c->stream3));
    c->cur = c->stream;
    // without shards the denominator comes out of the reduction: the backward waits for it here;
    // with shards it was exchanged at the start of the step and only the loss is joined, at the end
    if (!comm) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_red_join, 0));
  }
  if (comm && cnt_aside) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_cnt, 0));
  if (c->gdx_pending) {
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_g_join, 0));
    c->gdx_pending = false;
  }
  if (bwd && c->saved_fold_step) {
    // the forward ran k_fold_step_w + the dx GEMM; `den` holds num_present
    RET_IF(c->saved_B != B, SMPLB_ESTATE, "internal: forward state lost");
    int rows = cdiv(B, 128) * 128;
    TRY(launch_pose_bwd(c, B, c->ws_theta, c->ws_Rs, c->ws_J, c->ws_A, c->ws_dA, 1, c->ws_dx, 2, rows, c->ws_rowscale, nullptr,
                        odb, odt, den, w_kp, odc));
  } else if (bwd) {
    if (!have_mesh && c->saved_fold && c->fold_warp_kernels) {
      // keypoint-only backward: d kp loss -> d joints, d cam, du, dA in one kernel, then the GEMM
      RET_IF(c->saved_B != B, SMPLB_ESTATE, "internal: forward state lost");
      int rows = cdiv(B, 128) * 128;
      TRY(launch_fold_bwd(c, B, c->ws_A, nullptr, c->ws_dkp, jbuf, dcam, w_kp, den, odc, c->ws_dA, c->ws_dx, 2));
      TRY(launch_pose_bwd(c, B, c->ws_theta, c->ws_Rs, c->ws_J, c->ws_A, c->ws_dA, 1, c->ws_dx, 2, rows, c->ws_rowscale,
                          nullptr, odb, odt));
    } else {
      // d kp loss -> d joints, d cam; scale w_kp / num_present (global count if overridden / exchanged)
      TRY(launch_proj_bwd(c, B, K, jbuf, dcam, c->ws_dkp, 0, 0.f, 0.f, w_kp, den, 0, c->ws_djoints, odc));
      const float *dverts = nullptr;
      if (have_mesh) {
        TRY(ensure_buf(c, &c->ws_dverts, (size_t)c->ws_batch * c->V3, false));
        // (the mesh loss left its gradient unfinished: (d_sil + integer sign sums) / (3 + V) is applied here)
        TRY(launch_proj_bwd(c, B, c->V, vbuf, dcam, c->ws_dsil, 1, img_size, img_size, w_mesh, nullptr, 1, c->ws_dverts,
                            odc, c->ws_silcnt, (float)(3 + c->V)));
        dverts = c->ws_dverts;
      }
      TRY(smpl_backward_dev(c, B, dverts, c->ws_djoints, nullptr, odb, odt));
    }
  }
  TRY(join_verts(c));
  if (comm && red_aside) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_red_join, 0));   // the loss
  int rc = st.finish();
  guard.ok = rc == 0;
  return rc;
}

extern "C" int smplb_step(smplb_ctx *c, int B, const float *beta, const float *theta, const float *cam,
                          const float *kp_gt, const float *points_xy, const int32_t *offsets, int P, float w_kp,
                          float w_mesh, float img_size, int64_t kp_count_override, float *verts, float *joints, float *Rs,
                          float *kp_pred, float *loss_parts, float *d_beta, float *d_theta, float *d_cam, int flags,
                          int mem) {
  return step_core(c, B, beta, theta, cam, kp_gt, points_xy, offsets, P, nullptr, 0, 0, w_kp, w_mesh, img_size, kp_count_override,
                   verts, joints, Rs, kp_pred, loss_parts, d_beta, d_theta, d_cam, flags, mem);
}

extern "C" int smplb_step_seg(smplb_ctx *c, int B, const float *beta, const float *theta, const float *cam,
                              const float *kp_gt, const float *seg, int H, int W, float w_kp, float w_mesh, float img_size,
                              int64_t kp_count_override, float *verts, float *joints, float *Rs, float *kp_pred,
                              float *loss_parts, float *d_beta, float *d_theta, float *d_cam, int flags, int mem) {
  RET_IF(!seg, SMPLB_EINVAL, "seg is NULL");
  return step_core(c, B, beta, theta, cam, kp_gt, nullptr, nullptr, 0, seg, H, W, w_kp, w_mesh, img_size, kp_count_override, verts,
                   joints, Rs, kp_pred, loss_parts, d_beta, d_theta, d_cam, flags, mem);
}

// ------------------------------------------------------------------------------ NCCL (dlopen)
// The only exchange on this path is a <=512-float sum (loss numerators, counts, the 428
// gradient-penalty column sums), so NCCL is loaded lazily and is not a link-time dependency.
typedef struct {
  char internal[128];
} nccl_uid_t;
typedef int (*fn_getuid)(nccl_uid_t *);
typedef int (*fn_initrank)(void **, int, nccl_uid_t, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_destroy)(void *);
typedef const char *(*fn_errstr)(int);
static struct {
  void *lib;
  fn_getuid getuid;
  fn_initrank initrank;
  fn_allreduce allreduce;
  fn_destroy destroy;
  fn_errstr errstr;
} g_nccl = {};

static int nccl_load() {
  if (g_nccl.lib) return 0;
  // SMPLB_NCCL_LIB names the library to use.  The Python facade sets it to the NCCL that ships with the process's
  // PyTorch (when there is one): whichever libnccl.so.2 is mapped first serves every later request for that soname,
  // and a PyTorch imported AFTER this call needs its own, newer one.
  const char *names[] = {getenv("SMPLB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    if (!n || !*n) continue;
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  RET_IF(!g_nccl.lib, SMPLB_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.getuid = (fn_getuid)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.initrank = (fn_initrank)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  RET_IF(!g_nccl.getuid || !g_nccl.initrank || !g_nccl.allreduce || !g_nccl.destroy, SMPLB_ENCCL,
         "libnccl is missing a required symbol");
  return 0;
}
#define NCCL_TRY(expr)                                                                              \
  do {                                                                                              \
    int _r = (expr);                                                                                \
    if (_r != 0) {                                                                                  \
      smplb_set_error("%s -> NCCL error %d (%s)", #expr, _r, g_nccl.errstr ? g_nccl.errstr(_r) : "?"); \
      return SMPLB_ENCCL;                                                                           \
    }                                                                                               \
  } while (0)

extern "C" int smplb_comm_unique_id(void *id128) {
  RET_IF(!id128, SMPLB_EINVAL, "null id buffer");
  TRY(nccl_load());
  NCCL_TRY(g_nccl.getuid((nccl_uid_t *)id128));
  return 0;
}

extern "C" int smplb_comm_init(smplb_ctx *c, int nranks, int rank, const void *id128) {
  CHECK_CTX(c);
  RET_IF(nranks < 1 || rank < 0 || rank >= nranks || !id128, SMPLB_EINVAL, "bad nranks/rank/id");
  TRY(nccl_load());
  nccl_uid_t id;
  memcpy(&id, id128, sizeof(id));
  NCCL_TRY(g_nccl.initrank(&c->nccl_comm, nranks, id, rank));
  c->nranks = nranks;
  c->rank = rank;
  return 0;
}

// In place on `stream`; does NOT go through CHECK_CTX (which resets c->cur to the main stream):
// smplb_step calls this while it launches on stream3.
static int allreduce_on(smplb_ctx *c, void *dev_buf, size_t count, int nccl_dtype, cudaStream_t stream) {
  RET_IF(!c->nccl_comm, SMPLB_ENCCL, "smplb_comm_init has not been called");
  NCCL_TRY(g_nccl.allreduce(dev_buf, dev_buf, count, nccl_dtype, /*ncclSum*/ 0, c->nccl_comm, stream));
  return 0;
}

// public: on the context's main stream
extern "C" int smplb_comm_allreduce_sum(smplb_ctx *c, float *dev_buf, int count) {
  CHECK_CTX(c);
  RET_IF(!dev_buf || count < 1, SMPLB_EINVAL, "null buffer or count < 1");
  if (c->nranks == 1 && !c->nccl_comm) return 0;
  return allreduce_on(c, dev_buf, (size_t)count, /*ncclFloat*/ 7, c->stream);
}

// ---- mailbox exchange (k_exchange.cu): the peers' mailboxes mapped into this process -----------
static int x_ensure_mbox(smplb_ctx *c) {
  if (c->x_mbox) return 0;
  TRY(exchange_preload());
  CUDA_TRY(cudaMalloc((void **)&c->x_mbox, X_MBOX_ENTRIES * sizeof(XEntry)));
  CUDA_TRY(cudaMemset(c->x_mbox, 0, X_MBOX_ENTRIES * sizeof(XEntry)));
  return 0;
}

static void x_detach(smplb_ctx *c) {
  for (int r = 0; r < X_MAXR; ++r) {
    if (c->x_peers[r] && c->x_ipc[r]) cudaIpcCloseMemHandle(c->x_peers[r]);
    c->x_peers[r] = nullptr;
    c->x_ipc[r] = false;
  }
  c->x_attached = false;
}

extern "C" int smplb_comm_p2p_export(smplb_ctx *c, void *handle64) {
  CHECK_CTX(c);
  RET_IF(!handle64, SMPLB_EINVAL, "null handle buffer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  TRY(x_ensure_mbox(c));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, c->x_mbox));
  memcpy(handle64, &h, 64);
  return 0;
}

extern "C" int smplb_comm_p2p_attach(smplb_ctx *c, int nranks, int rank, const void *handles) {
  CHECK_CTX(c);
  RET_IF(nranks < 1 || nranks > X_MAXR || rank < 0 || rank >= nranks || !handles, SMPLB_EINVAL,
         "1 <= nranks <= %d, 0 <= rank < nranks and non-null handles required", X_MAXR);
  RET_IF(c->nccl_comm && (c->nranks != nranks || c->rank != rank), SMPLB_EINVAL,
         "nranks / rank differ from the attached NCCL communicator's");
  TRY(x_ensure_mbox(c));
  x_detach(c);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      c->x_peers[r] = c->x_mbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)r * 64, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      x_detach(c);
      smplb_set_error("cudaIpcOpenMemHandle(rank %d) -> %s (no peer access between the GPUs?)", r, cudaGetErrorString(e));
      return SMPLB_ECUDA;
    }
    c->x_peers[r] = (XEntry *)p;
    c->x_ipc[r] = true;
  }
  c->nranks = nranks;
  c->rank = rank;
  c->x_epoch = 0;
  c->x_attached = true;
  return 0;
}

// Same, for ranks that are contexts of ONE process (tests; several GPUs driven by one process).
extern "C" int smplb_comm_p2p_attach_local(smplb_ctx *c, int nranks, int rank, smplb_ctx *const *peers) {
  CHECK_CTX(c);
  RET_IF(nranks < 1 || nranks > X_MAXR || rank < 0 || rank >= nranks || !peers, SMPLB_EINVAL,
         "1 <= nranks <= %d, 0 <= rank < nranks and non-null peers required", X_MAXR);
  RET_IF(peers[rank] != c, SMPLB_EINVAL, "peers[rank] must be the context itself");
  TRY(x_ensure_mbox(c));
  x_detach(c);
  for (int r = 0; r < nranks; ++r) {
    smplb_ctx *p = peers[r];
    RET_IF(!p, SMPLB_EINVAL, "peers[%d] is NULL", r);
    if (p != c) {
      CUDA_TRY(cudaSetDevice(p->device));
      int rc = x_ensure_mbox(p);
      CUDA_TRY(cudaSetDevice(c->device));
      if (rc) return rc;
      if (p->device != c->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) {
          smplb_set_error("cudaDeviceEnablePeerAccess(%d -> %d) -> %s", c->device, p->device, cudaGetErrorString(e));
          return SMPLB_ECUDA;
        }
      }
    }
    c->x_peers[r] = p->x_mbox;
  }
  c->nranks = nranks;
  c->rank = rank;
  c->x_epoch = 0;
  c->x_attached = true;
  return 0;
}

// Test hook (csrc/smplb_debug.h): writes rank `from_rank`'s entry of `epoch` into THIS context's mailbox, as that
// rank's push would.  Lets a single-GPU test run the ranks one after the other: no kernel ever waits for a kernel
// that has not been launched yet (two spinning kernels of one GPU are not guaranteed to run concurrently).
extern "C" int smplb_debug_p2p_inject(smplb_ctx *c, int kind, unsigned epoch, int from_rank, float v0, float v1,
                                      long long cnt) {
  CHECK_CTX(c);
  RET_IF(!c->x_mbox || kind < 0 || kind > 1 || from_rank < 0 || from_rank >= X_MAXR, SMPLB_EINVAL, "no mailbox / bad arguments");
  XEntry e = {};
  e.v[0] = v0;
  e.v[1] = v1;
  e.cnt = cnt;
  e.flag = epoch;
  XEntry *dst = c->x_mbox + ((size_t)kind * X_SLOTS + (epoch % X_SLOTS)) * X_MAXR + from_rank;
  CUDA_TRY(cudaMemcpyAsync(dst, &e, sizeof(e), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

// 0 = every exchange so far completed; 1 = a pull timed out (a peer never arrived): the step that
// saw it returned a NaN loss and zero gradients' denominator.
extern "C" int smplb_comm_status(smplb_ctx *c, int *status) {
  CHECK_CTX(c);
  RET_IF(!status, SMPLB_EINVAL, "null status");
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(status, c->x_status, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int smplb_comm_destroy(smplb_ctx *c) {
  if (!c) return 0;
  if (c->x_attached) {
    cudaSetDevice(c->device);
    x_detach(c);
  }
  if (!c->nccl_comm) {
    c->nranks = 1;
    c->rank = 0;
    return 0;
  }
  if (g_nccl.destroy) g_nccl.destroy(c->nccl_comm);
  c->nccl_comm = nullptr;
  c->nranks = 1;
  c->rank = 0;
  return 0;
}
