// The path's one exchange between batch shards, fused into the loss kernels.
//
// A step on N ranks needs two tiny sums over the ranks: the visibility count 2 * #{vis != 0}
// (kp_reprojection_loss, src/ops.py:35-47: the denominator of the loss and of every keypoint
// gradient) and the loss numerators {sum vis * |gt - pred|, mesh sum}.  The count depends on
// kp_gt alone, so it is exchanged at the START of the step and the backward never waits for a
// peer; the numerators only feed the reported loss and are exchanged next to the backward.
//
// Transport: every context owns a mailbox in its GPU's memory that the peers map (CUDA IPC
// between processes, plain pointers inside one process).  A rank PUSHES its partial into slot
// [kind][epoch % X_SLOTS][rank] of every peer's mailbox over NVLink (data, a system-scope fence,
// then the epoch as the flag) and PULLS by spinning on the flags of its own mailbox -- local
// memory -- and adding the entries in rank order, so every rank forms the same bits.  The push
// and the pull sit inside the kernels that produce / consume the values (k_count_exchange,
// k_reduce_exchange_finalize): no collective launch, no host round trip.  NCCL stays available
// as the second backend (smplb_comm_init) and cross-checks this one in the tests.
//
// Slot reuse: a rank can only push epoch e + 2 after it pulled epoch e + 1, which needs every
// peer's push of e + 1, which those peers issue after their own pull of e (stream order inside
// a context) -- so two slots would do; there are four.
// Ordering contract (same as NCCL's for several communicators): all ranks issue the steps of
// their contexts in the same order.
#include "smplb_internal.h"

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Thread r < nranks: writes this rank's entry into peer r's mailbox.
__device__ __forceinline__ void x_push(const XArgs &x, int kind, int r, float v0, float v1, long long cnt) {
  XEntry *e = x.peers[r] + ((size_t)kind * X_SLOTS + (x.epoch % X_SLOTS)) * X_MAXR + x.rank;
  volatile float *vf = e->v;
  vf[0] = v0;
  vf[1] = v1;
  *(volatile long long *)&e->cnt = cnt;
  __threadfence_system();
  st_release_sys(&e->flag, x.epoch);
}

// Thread r < nranks: waits for rank r's entry of this epoch in the own mailbox.  A peer that
// never arrives (crashed rank) must not hang the GPU: after x.timeout_ns the wait gives up and
// the caller poisons its outputs (NaN loss, zero count).
__device__ __forceinline__ bool x_pull(const XArgs &x, int kind, int r, float *v0, float *v1, long long *cnt) {
  const XEntry *e = x.peers[x.rank] + ((size_t)kind * X_SLOTS + (x.epoch % X_SLOTS)) * X_MAXR + r;
  unsigned long long t0 = global_ns();
  unsigned spins = 0;
  while (ld_acquire_sys(&e->flag) != x.epoch) {
    __nanosleep(40);
    if ((++spins & 1023u) == 0 && global_ns() - t0 > x.timeout_ns) return false;
  }
  const volatile float *vf = e->v;
  *v0 = vf[0];
  *v1 = vf[1];
  *cnt = *(const volatile long long *)&e->cnt;
  return true;
}

// ---- step start: visibility count of the shard, exchanged ------------------------------------
// den_out receives the denominator every keypoint gradient of this step divides by: the override
// if > 0, else the count summed over the ranks (mode 1: mailbox; mode 2: the LOCAL count, which
// the host all-reduces in place with NCCL right behind this kernel; mode 0: local).
#define XC_THREADS 256
__global__ void __launch_bounds__(XC_THREADS) k_count_exchange(int n, const float *__restrict__ kp_gt,
                                                               long long count_override, int mode, XArgs x,
                                                               long long *__restrict__ den_out,
                                                               int *__restrict__ status) {
  __shared__ long long redc[XC_THREADS];
  __shared__ long long rank_cnt[X_MAXR];
  __shared__ int ok[X_MAXR];
  const int t = threadIdx.x;
  long long c = 0;
  if (count_override <= 0)
    for (int i = t; i < n; i += XC_THREADS) c += (kp_gt[3 * (size_t)i + 2] != 0.0f) ? 2 : 0;
  redc[t] = c;
  __syncthreads();
  for (int o = XC_THREADS / 2; o > 0; o >>= 1) {
    if (t < o) redc[t] += redc[t + o];
    __syncthreads();
  }
  if (count_override > 0) {
    if (t == 0) *den_out = count_override;
    return;
  }
  if (mode != 1) {
    if (t == 0) *den_out = redc[0];
    return;
  }
  if (t < x.nranks) x_push(x, 0, t, 0.f, 0.f, redc[0]);
  if (t < x.nranks) {
    float a, b;
    long long rc = 0;
    ok[t] = x_pull(x, 0, t, &a, &b, &rc) ? 1 : 0;
    rank_cnt[t] = rc;
  }
  __syncthreads();
  if (t == 0) {
    long long s = 0;
    bool good = true;
    for (int r = 0; r < x.nranks; ++r) {
      s += rank_cnt[r];
      good = good && ok[r];
    }
    *den_out = good ? s : 0;
    if (!good) *status = 1;
  }
}

// ---- after k_fold_step_w (or the mesh loss): numerators reduced, exchanged, loss finalized -------
// part / cnt: per-body partials (NULL: scal[0] already holds the local kp numerator);
// scal = {kp numerator, local count as float, mesh sum}: overwritten with the GLOBAL values.
// den: the denominator k_count_exchange left (global).  out = loss_parts[4].
#define XR_THREADS 256
__global__ void __launch_bounds__(XR_THREADS) k_reduce_exchange_finalize(int B, const float *__restrict__ part,
                                                                         float w_kp, float w_mesh, int have_mesh, XArgs x,
                                                                         float *__restrict__ scal,
                                                                         const long long *__restrict__ den,
                                                                         float *__restrict__ out,
                                                                         int *__restrict__ status) {
  __shared__ float red[XR_THREADS];
  __shared__ float rv0[X_MAXR], rv1[X_MAXR];
  __shared__ int ok[X_MAXR];
  const int t = threadIdx.x;
  if (part) {
    float s = 0.f;
    for (int i = t; i < B; i += XR_THREADS) s += part[i];
    red[t] = s;
    __syncthreads();
    for (int o = XR_THREADS / 2; o > 0; o >>= 1) {
      if (t < o) red[t] += red[t + o];
      __syncthreads();
    }
  } else {
    if (t == 0) red[0] = scal[0];
    __syncthreads();
  }
  const float num = red[0];
  const float mesh = have_mesh ? scal[2] : 0.0f;
  if (t < x.nranks) x_push(x, 1, t, num, mesh, 0);
  if (t < x.nranks) {
    long long dummy;
    ok[t] = x_pull(x, 1, t, &rv0[t], &rv1[t], &dummy) ? 1 : 0;
  }
  __syncthreads();
  if (t == 0) {
    float s0 = 0.f, s1 = 0.f;
    bool good = true;
    for (int r = 0; r < x.nranks; ++r) {   // rank order: the same bits on every rank
      s0 += rv0[r];
      s1 += rv1[r];
      good = good && ok[r];
    }
    long long d = *den;
    float kp = d > 0 ? s0 / (float)d : 0.0f;
    scal[0] = s0;
    scal[1] = (float)d;
    scal[2] = s1;
    out[0] = s0;
    out[1] = (float)d;
    out[2] = s1;
    out[3] = good ? w_kp * kp + w_mesh * s1 : __int_as_float(0x7fc00000);
    if (!good) *status = 1;
  }
}

// NCCL backend / local: scal = {numerator, -, mesh} (all-reduced by the host call in between when
// a communicator is attached), den from k_count_exchange.
__global__ void k_finalize_den(float w_kp, float w_mesh, int have_mesh, const float *__restrict__ scal,
                               const long long *__restrict__ den, float *__restrict__ out) {
  long long d = *den;
  float kp = d > 0 ? scal[0] / (float)d : 0.0f;
  float ml = have_mesh ? scal[2] : 0.0f;
  out[0] = scal[0];
  out[1] = (float)d;
  out[2] = ml;
  out[3] = w_kp * kp + w_mesh * ml;
}

// With lazy module loading the first launch of a kernel loads it, which can wait for the device to
// drain -- fatal when the kernel running is a rank of the SAME process waiting for that very launch.
// Attaching therefore loads the exchange kernels up front.
int exchange_preload() {
  cudaFuncAttributes a;
  CUDA_TRY(cudaFuncGetAttributes(&a, k_count_exchange));
  CUDA_TRY(cudaFuncGetAttributes(&a, k_reduce_exchange_finalize));
  CUDA_TRY(cudaFuncGetAttributes(&a, k_finalize_den));
  return 0;
}

static XArgs make_xargs(smplb_ctx *c) {
  XArgs x;
  for (int r = 0; r < X_MAXR; ++r) x.peers[r] = r < c->nranks ? c->x_peers[r] : nullptr;
  x.nranks = c->nranks;
  x.rank = c->rank;
  x.epoch = c->x_epoch;
  x.timeout_ns = c->x_timeout_ns;
  return x;
}

int launch_count_exchange(smplb_ctx *c, int B, const float *kp_gt, long long count_override, int mode, long long *den) {
  XArgs x = {};
  if (mode == 1) x = make_xargs(c);
  LAUNCH(c, "kp_count_exchange", 1, XC_THREADS, 0, k_count_exchange, B * c->K, kp_gt, count_override, mode, x, den,
         c->x_status);
  return 0;
}

int launch_reduce_exchange_finalize(smplb_ctx *c, int B, const float *part, float w_kp, float w_mesh, int have_mesh,
                                    const long long *den, float *loss_parts) {
  XArgs x = make_xargs(c);
  LAUNCH(c, "reduce_exchange_finalize", 1, XR_THREADS, 0, k_reduce_exchange_finalize, B, part, w_kp, w_mesh, have_mesh, x,
         c->ws_scal, den, loss_parts, c->x_status);
  return 0;
}

int launch_finalize_den(smplb_ctx *c, float w_kp, float w_mesh, int have_mesh, const long long *den, float *loss_parts) {
  LAUNCH(c, "finalize_loss", 1, 1, 0, k_finalize_den, w_kp, w_mesh, have_mesh, c->ws_scal, den, loss_parts);
  return 0;
}
