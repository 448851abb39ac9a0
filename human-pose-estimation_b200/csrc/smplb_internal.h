// Internal declarations shared by the kernel translation units of libsmplb.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "smplb.h"
#include "smplb_debug.h"

#define NJ 24          // SMPL joints
#define NPF 207        // pose-feature length, 23 * 9
#define KX 224         // padded length of x_ext = [pose_feature(207) | beta(NB) | 1 | 0...]
#define VSPLIT 4       // vertex-range splits of the skinning backward at large batches (fixed-order partial sums)
#define VSPLIT_MAX 16  // ... and at small ones: CTAs = (B / 16) x splits, so B = 1024 gets 16 splits (1024 CTAs instead of 256)
static inline int skin_bwd_splits(int B) {
  if (B >= 4096) return VSPLIT;
  int s = 16384 / (B > 0 ? B : 1);
  return s < VSPLIT ? VSPLIT : (s > VSPLIT_MAX ? VSPLIT_MAX : s);
}
static inline size_t skin_bwd_part_rows(int max_batch) {   // rows of [24 x 12] the partial-sum buffer needs for any B <= max_batch
  size_t a = (size_t)VSPLIT * max_batch, b = (size_t)(max_batch < 1024 ? (size_t)VSPLIT_MAX * max_batch : 16384);
  return a > b ? a : b;
}
#define MAXK SMPLB_MAX_KEYPOINTS

struct Tree {
  signed char parent[NJ];
  signed char depth[NJ];
  int max_depth;
  // for the subtree-sum backward (k_pose_bwd_reg): the first three children of each joint in index order (-1: none),
  // per depth d the largest child count among the joints of depth d - 1 (2 bits each), and the largest child count
  signed char child[NJ][3];
  unsigned level_slots;
  int max_children;
};

// ---- mailbox exchange between batch shards (k_exchange.cu)
#define X_SLOTS 4
#define X_MAXR 16
struct __align__(16) XEntry {
  float v[2];
  long long cnt;
  unsigned flag;   // epoch of the entry
  unsigned pad[3];
};
#define X_MBOX_ENTRIES (2 * X_SLOTS * X_MAXR)   // [kind: 0 counts, 1 numerators][slot][rank]
struct XArgs {
  XEntry *peers[X_MAXR];   // mailbox of every rank (own one included), mapped into this process
  int nranks, rank;
  unsigned epoch;
  unsigned long long timeout_ns;
};

struct ProfRec {
  const char *name;
  cudaEvent_t e0, e1;
};

struct smplb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;    // main stream: everything the caller orders against
  cudaStream_t stream2 = nullptr;   // side stream: the 6890-vertex blend + skinning, overlapped with the keypoint path
  cudaStream_t cur = nullptr;       // stream the LAUNCH macro uses
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // third stream: the loss reduction (+ the all-reduce over ranks) of a keypoint step runs next to the dx GEMM
  cudaStream_t stream3 = nullptr;
  cudaEvent_t ev_red_fork = nullptr, ev_red_join = nullptr;
  bool red_fork_recorded = false;   // ev_red_fork was recorded behind k_fold_step_w in this step
  // The GPU dispatches the grids of equal-priority streams in launch order, and a grid that cannot be
  // placed blocks every later one, whatever its stream: the two tcgen05 GEMMs of the keypoint path
  // wait for SMs without a resident vertex-kernel CTA (tensor memory, shared memory), and while one
  // of them waits nothing launched after it starts -- not even the small per-body kernels of other
  // contexts that would fit next to the vertex kernel.  So the GEMMs (and the vertex kernel) run on
  // low-priority streams and everything else on high-priority ones, whose grids are dispatched first.
  cudaStream_t stream_g = nullptr;  // low priority: fold_gemm_u, fold_gemm_dx
  cudaEvent_t ev_g_fork = nullptr, ev_g_join = nullptr;
  bool gdx_pending = false;         // fold_gemm_dx is on stream_g and not yet joined into the main stream
  int use_prio = 1;                 // smplb_debug_set("prio", 0): the GEMMs stay on the main stream
  bool verts_pending = false;       // stream2 work not yet joined into the main stream
  int use_overlap = 1;              // smplb_debug_set("overlap", 0) keeps everything on the main stream
  int V = 0, NB = 0, K = 0, max_batch = 0;
  int V3 = 0;       // 3V
  int Vp = 0;       // V rounded up to 128
  int pitch = 0;    // 3 * Vp: row pitch (floats) of v_posed / dp / Dext.  These are PLANAR: column c * Vp + v holds
                    // coordinate c of vertex v (so a warp reads 32 vertices of one coordinate as one 128 B line)
  int ksplit = 1;   // split-K factor of the blend backward GEMM
  Tree tree;
  // ---- constants on the device
  float *d_vt = nullptr, *d_shapedirs = nullptr, *d_posedirs = nullptr;
  float *d_W = nullptr, *d_JR = nullptr;
  float *d_Dext = nullptr;   // [KX][pitch]: rows 0..206 posedirs, 207..207+NB-1 shapedirs, 207+NB v_template
  float *d_J0 = nullptr;     // [24*3]      J_regressor^T v_template
  float *d_Jdirs = nullptr;  // [24*3][NB]  J_regressor^T shapedirs
  int *d_kcsr_off = nullptr, *d_kcsr_idx = nullptr;  // joint_regressor by keypoint (forward gather)
  float *d_kcsr_val = nullptr;
  int *d_vcsr_off = nullptr, *d_vcsr_k = nullptr;    // joint_regressor by vertex (backward)
  float *d_vcsr_val = nullptr;
  // ---- active vertices: rows of joint_regressor with a non-zero.  With no upstream d_verts
  //      the backward only has to walk these (every other vertex has a zero gradient).
  int n_act = 0, Vpa = 0, pitch_act = 0;   // compact planar layout: column c * Vpa + a
  int use_compact = 1;             // smplb_debug_set("compact_bwd", 0) forces the dense walk (validation)
  int *d_act_idx = nullptr;
  float *d_act_W = nullptr;        // [n_act][24]
  int *d_acsr_off = nullptr, *d_acsr_k = nullptr;
  float *d_acsr_val = nullptr;
  float *d_Dext_act = nullptr;     // [KX][pitch_act]
  float *ws_dp_act = nullptr;      // [B][pitch_act]
  // compact keypoint path: the same blend / skinning kernels run on the active vertices only, so
  // joints, the keypoint loss and its backward never touch the 6890-vertex tensors
  void *d_Dt16_act = nullptr;      // [3*Vpa][256] fp16: rows of Dt16 gathered
  void *d_W16_act = nullptr;       // [Vpa][64] fp16
  alignas(64) unsigned char map_d_act[128];
  alignas(64) unsigned char map_w_act[128];
  int *d_kcsr_slot = nullptr;      // kcsr_idx re-indexed to active slots
  float *ws_vposed_act = nullptr;  // [B][3*Vpa]
  float *ws_verts_act = nullptr;   // [B][n_act][3]
  bool compact_ok = false;
  int l2_chunk = 0;                // samples per L2-resident blend->skin chunk (0 = off); smplb_debug_set("l2_chunk", n)
  bool saved_full = false;         // ws_vposed holds the last forward's full v_posed
  bool saved_compact = false;      // ws_vposed_act holds the last forward's compact v_posed
  float *saved_verts = nullptr;    // where the last forward wrote verts (caller's buffer or ws_verts)
  // ---- folded keypoint path (k_fold.cu): joints and their backward from x and A alone
  bool fold_ok = false;
  int use_fold = 1;                // smplb_debug_set("fold", 0) falls back to the per-vertex keypoint path
  int fold_nu = 0, fold_nup = 0;   // 3 * 24 * K and its round-up to 128
  float *d_G = nullptr, *d_cc = nullptr;
  void *d_G16 = nullptr, *d_Gt16 = nullptr;
  alignas(64) unsigned char map_g1[128];
  alignas(64) unsigned char map_g2[128];
  float fold_scale = 1.f, fold_inv_scale = 1.f;
  float *ws_U = nullptr;           // [B][fold_nup]
  void *ws_du16 = nullptr;         // [B][3 * fold_nup] fp16
  float *ws_rowscale = nullptr;    // [B]
  void *ws_x16b = nullptr;         // unused since the fold GEMM reads the blend operand rows (ws_x16); kept for the kernel signature
  bool saved_fold = false;
  bool saved_fold_step = false;    // the forward already ran the fused keypoint forward + backward (k_fold_step_w)
  int use_fold_step = 1;           // smplb_debug_set("fold_step", 0): separate forward / backward kernels
  int use_pose_bwd_reg = 1;        // smplb_debug_set("pose_bwd_reg", 0): the serial reverse walk in shared memory (validation)
  int fold_warp_kernels = 1;       // smplb_debug_set("fold_warp", 0): CTA-per-body reference kernels
  // ---- tcgen05 blend path (k_blend_tc.cu)
  bool tc_ok = false;          // operands built, tensor map encoded
  int use_tc = 1;              // smplb_debug_set("blend_tc", 0) selects the FP32 CUDA-core GEMM (validation)
  float tc_scale = 1.f, tc_inv_scale = 1.f;
  void *d_Dt16 = nullptr;      // [pitch][256] fp16, K-major
  alignas(64) unsigned char map_d[128];   // CUtensorMap of Dt16
  alignas(64) unsigned char map_d32[128]; // the same rows as [128 x 16]-column boxes, SWIZZLE_32B (k_body_res.cu: last k-block)
  int num_sms = 148;
  void *ws_x16 = nullptr;      // [B][256] fp16 operand rows
  // ---- tcgen05 skinning path (k_skin_tc.cu)
  bool skin_tc_ok = false;
  int use_skin_tc = 1;         // smplb_debug_set("skin_tc", 0) selects the FP32 CUDA-core skinning kernel
  // ---- tcgen05 skinning backward (k_skin_bwd_tc.cu)
  bool skin_bwd_tc_ok = false;
  int use_skin_bwd_tc = 1;     // smplb_debug_set("skin_bwd_tc", 0) selects the FP32 CUDA-core kernel k_skin_bwd for the dense walk
  void *d_WT16 = nullptr;      // [64][Vp] bf16: W^T hi (rows 0..23) and lo (rows 32..55)
  alignas(64) unsigned char map_wt[128];   // CUtensorMap of WT16
  void *d_W16 = nullptr;       // [Vp][64] fp16, 16-column windows (k_skin_tc.cu)
  alignas(64) unsigned char map_w[128];   // CUtensorMap of W16
  void *ws_A16 = nullptr;      // [B*12][64] fp16, row (b, 4r+d), 16-column windows (k_skin_tc.cu)
  // ---- fused blend + skinning (k_body_tc.cu): verts without the v_posed round trip
  bool body_tc_ok = false;
  int body_pairs = 0;          // smplb_debug_set("body_pairs", n): CTA pairs of the vertex kernel (0 = automatic, -1 = one per SM pair)
  int pairs_auto = 0;          // set per launch by smpl_forward_dev: pairs to use when body_pairs == 0 (0 = one per SM pair)
  int use_fused = 1;           // smplb_debug_set("fused", 0) selects the two-kernel path (which saves v_posed)
  // ---- workspace, sized for max_batch (grown on demand)
  int ws_batch = 0;
  float *ws_x = nullptr;       // [B][KX]
  float *ws_Rs = nullptr;      // [B][24][9]
  float *ws_J = nullptr;       // [B][24][3]
  float *ws_A = nullptr;       // [B][24][12]  rows (R r0 r1 r2 | t)
  float *ws_Jtr = nullptr;     // [B][24][3]
  float *ws_vposed = nullptr;  // [B][pitch]
  float *ws_verts = nullptr;   // [B][3V]   used when the caller does not ask for verts
  float *ws_joints = nullptr;  // [B][K][3]
  float *ws_kp = nullptr;      // [B][K][2]
  float *ws_dkp = nullptr;     // [B][K][2]  unscaled d loss / d kp_pred
  float *ws_djoints = nullptr; // [B][K][3]
  float *ws_dverts = nullptr;  // [B][3V]
  float *ws_silpred = nullptr; // [B][V][2]
  float *ws_dsil = nullptr;    // [B][V][2]
  int *ws_silcnt = nullptr;    // [B][V][2] integer sign sums of the pixel->vertex term
  float *ws_dp = nullptr;      // [B][pitch]
  void *ws_dp16 = nullptr;     // [B][3 * pitch] bf16: hi | lo | hi (operand of the tcgen05 blend-transpose GEMM)
  void *d_Dbf = nullptr;       // [KX][3 * pitch] bf16: Dext hi | hi | lo
  alignas(64) unsigned char map_dbf[128];
  bool blend_bwd_tc_ok = false;
  int use_blend_bwd_tc = 1;    // smplb_debug_set("blend_bwd_tc", 0): FP32 CUDA-core GEMM (cross-check)
  float *ws_dA = nullptr;      // [skin_bwd_splits(B)][B][288]
  float *ws_dx = nullptr;      // [ksplit][B][KX]
  float *ws_part = nullptr;    // per-body / per-block float partials
  int *ws_cnt = nullptr;       // per-body int partials
  float *ws_scal = nullptr;    // small device scalars: [0]=kp abs_sum [2]=mesh [3]=total
  long long *ws_cnt64 = nullptr;  // [0] = kp num_present
  float *ws_theta = nullptr, *ws_beta = nullptr;   // copies kept for backward
  float *ws_gp = nullptr;      // gradient-penalty partials
  unsigned int *ws_ticket = nullptr;   // last-CTA-done counter of k_critic_gp
  size_t ws_gp_cap = 0;        // floats
  float *ws_mesh_part = nullptr;   // mesh-loss per-CTA partials
  float *ws_vdist = nullptr;       // [B][V] vertex->pixel distances (summed per image in index order)
  size_t ws_vdist_cap = 0;
  void *ws_grid = nullptr;         // uniform-grid search workspace (k_loss.cu)
  size_t ws_grid_cap = 0;
  float *ws_segpts = nullptr;      // smplb_step_seg: the compacted silhouette points [P][2] ...
  int *ws_segoff = nullptr;        // ... and their offsets [B + 1]
  size_t ws_segpts_cap = 0, ws_segoff_cap = 0;
  unsigned *ws_segbits = nullptr;  // where(seg > 0) as a bitmap, written by k_sil_count for the fill [B][ceil(HW / 32)]
  size_t ws_segbits_cap = 0;
  int use_mesh_lattice = 1;        // smplb_debug_set("mesh_lattice", 0): vertex -> pixel search through the binned grid for every image
  int use_mesh_grid = 1;           // smplb_debug_set("mesh_grid", 0): brute-force scan (the reference's own algorithm)
  size_t ws_mesh_part_cap = 0;
  size_t ws_mesh_cap = 0;      // elements of ws_silpred / ws_dsil / ws_silcnt
  int saved_B = 0;
  unsigned attr_done = 0;   // bit i: cudaFuncSetAttribute done for kernel i on this device
  // ---- misc
  void *flush_buf = nullptr;
  size_t flush_bytes = 0;
  cudaEvent_t timer0[16] = {}, timer1[16] = {};
  int64_t launches = 0;
  bool profile = false;
  bool profile_serial = false;
  bool profile_trace = false;      // smplb_profile_enable(ctx, 2): per-launch timeline instead of per-kernel sums
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  // NCCL (dlopen'ed)
  void *nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  int comm_backend = 0;            // smplb_debug_set("comm_backend", 1): NCCL even when the mailboxes are attached
  // mailbox exchange (k_exchange.cu): own mailbox + the peers' mapped ones
  XEntry *x_mbox = nullptr;
  XEntry *x_peers[X_MAXR] = {};
  bool x_ipc[X_MAXR] = {};         // x_peers[r] came from cudaIpcOpenMemHandle
  bool x_attached = false;
  unsigned x_epoch = 0;            // exchanges issued so far (the same on every rank)
  unsigned long long x_timeout_ns = 20ull * 1000 * 1000 * 1000;
  int *x_status = nullptr;         // device flag: 1 after a pull timed out
  cudaEvent_t ev_step0 = nullptr, ev_cnt = nullptr;   // step start -> stream3; count exchanged -> main stream
};

void smplb_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      smplb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
      return SMPLB_ECUDA;                                                                       \
    }                                                                                           \
  } while (0)

#define RET_IF(cond, code, ...)   \
  do {                            \
    if (cond) {                   \
      smplb_set_error(__VA_ARGS__); \
      return (code);              \
    }                             \
  } while (0)

#define TRY(expr)          \
  do {                     \
    int _r = (expr);       \
    if (_r != 0) return _r; \
  } while (0)

// NVTX range named after the reference's tf.name_scope of the same stage, so an
// nsys / ncu timeline of a trainer reads like the reference's graph.  Free when no tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

struct ProfScope {
  smplb_ctx *c;
  ProfRec r;
  ProfScope(smplb_ctx *ctx, const char *name);
  ~ProfScope();
};

// Launch helper: counts launches, optional per-kernel event timing, checks the launch.
#define LAUNCH(ctx, name, grid, block, smem, kernel, ...)                                   \
  do {                                                                                      \
    {                                                                                       \
      ProfScope _ps((ctx), (name));                                                         \
      kernel<<<(grid), (block), (smem), (ctx)->cur>>>(__VA_ARGS__);                         \
    }                                                                                       \
    (ctx)->launches++;                                                                      \
    cudaError_t _le = cudaGetLastError();                                                   \
    if (_le != cudaSuccess) {                                                               \
      smplb_set_error("launch %s failed: %s", (name), cudaGetErrorString(_le));             \
      return SMPLB_ECUDA;                                                                   \
    }                                                                                       \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
#ifdef __CUDACC__
__device__ __forceinline__ int cdiv_dev(int a, int b) { return (a + b - 1) / b; }
#endif

// ---- kernel launchers (device pointers only), one per stage -------------------------------
// k_pose.cu
int launch_pose_fwd(smplb_ctx *c, int B, const float *beta, const float *theta, float *Rs, float *J, float *A,
                    float *Jtr, float *x, void *x16, void *A16, void *x16b);
int launch_pose_bwd(smplb_ctx *c, int B, const float *theta, const float *Rs, const float *J, const float *A,
                    const float *dA_part, int n_dA_parts, const float *dx_part, int ksplit, int dx_rows,
                    const float *rowscale, const float *d_Rs, float *d_beta, float *d_theta,
                    const long long *den = nullptr, float gscale = 1.0f, float *d_cam = nullptr);
int launch_rodrigues(smplb_ctx *c, int N, const float *theta, float *R);
int launch_global_rigid(smplb_ctx *c, int B, const float *Rs, const float *Js, float *new_J, float *A44);
int launch_skew(smplb_ctx *c, int N, const float *vec, float *out);
int launch_lrotmin(smplb_ctx *c, int B, const float *theta, float *out);
// k_blend.cu
int launch_blend_fwd(smplb_ctx *c, int B, const float *x, float *v_posed);
int launch_blend_bwd(smplb_ctx *c, int B, const float *dp, float *dx_part, bool compact, int ksplit);
// tcgen05 version of the dense one: dp16 = [B][3 * pitch] bf16 (hi | lo | hi, written by k_skin_bwd) against
// Dext as bf16 (hi | hi | lo); returns the split-K factor and row pitch of dx_part through the pointers
int blend_bwd_tc_init(smplb_ctx *c);
int launch_blend_bwd_tc(smplb_ctx *c, int B, const void *dp16, float *dx_part, int *ksplit, int *dx_rows);
// k_blend_tc.cu
int blend_tc_init(smplb_ctx *c);
int launch_blend_fwd_tc(smplb_ctx *c, int B, const void *x16, float *v_posed, bool act);
// k_fold.cu
int fold_init(smplb_ctx *c);
int launch_fold_gemm_u(smplb_ctx *c, int B, const void *x16b);
int launch_fold_fwd(smplb_ctx *c, int B, const void *x16b, const float *A, const float *cam, const float *kp_gt,
                    float *joints, float *kp_pred, float *dkp, float *part, int *cnt);
int launch_fold_bwd(smplb_ctx *c, int B, const float *A, const float *d_joints, const float *dkp, const float *joints,
                    const float *cam, float gscale, const long long *den, float *d_cam, float *dA_part, float *dx_part,
                    int ksplit);
int launch_fold_step(smplb_ctx *c, int B, const float *A, const float *cam, const float *kp_gt, float *joints,
                     float *kp_pred, float *part, int *cnt, float *d_cam, float *dA_part, float *dx_part, int ksplit);
int launch_reduce_finalize(smplb_ctx *c, int B, float w_kp, float w_mesh, long long count_override, float *loss_parts);
// k_mesh_lattice.cu
size_t mesh_lattice_workspace(int B);
int launch_mesh_lattice_search(smplb_ctx *c, int B, int V, const int *offsets, const void *ws, const float *gparam,
                               const float4 *sortedB, float *vdist, float *d_sil, int *ind_ba);
// k_gemm_tc.cu
int tc_make_map(void *map, int is_f32, const void *ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                uint32_t box_inner, uint32_t box_outer, int swizzle = 1);
int launch_gemm_tc(smplb_ctx *c, const char *name, int M, int N, int K, const void *A16, const void *map_b, float *C,
                   int ldc, int ksplit, float scale, int bf16 = 0);
// k_skin_bwd_tc.cu
int skin_bwd_tc_init(smplb_ctx *c);
int launch_skin_bwd_tc(smplb_ctx *c, int B, const void *A16, const float *v_posed, const float *d_verts, const float *d_joints,
                       float *dp, void *dp16, float *dA_part, int *n_parts);
// k_skin_tc.cu
int skin_tc_init(smplb_ctx *c);
int launch_skin_fwd_tc(smplb_ctx *c, int B, const void *A16, const float *v_posed, float *verts, bool act);
int compact_tc_init(smplb_ctx *c);
// k_body_tc.cu
int body_tc_init(smplb_ctx *c);
int launch_body_fwd_tc(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts);
// k_body_pair.cu (cta_group::2 variant; needs an even number of 128-vertex tiles)
int body_pair_init(smplb_ctx *c);
int launch_body_fwd_pair(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts);
// k_body_res.cu (cta_group::2 with the Dt16 tile resident in shared memory and W16 in tensor memory)
int body_res_init(smplb_ctx *c);
int launch_body_fwd_res(smplb_ctx *c, int B, const void *x16, const void *A16, float *verts, int variant);
// k_skin.cu
int launch_skin_fwd(smplb_ctx *c, int B, const float *A, const float *v_posed, float *verts);
int launch_joints(smplb_ctx *c, int B, const float *verts, const float *cam, const float *kp_gt, float *joints,
                  float *kp_pred, float *dkp, float *part, int *cnt, bool act);
int launch_skin_bwd(smplb_ctx *c, int B, const float *A, const float *v_posed, const float *d_verts,
                    const float *d_joints, float *dp, float *dA_part, int mode, void *dp16 = nullptr);
int launch_proj(smplb_ctx *c, int B, int N, const float *X, const float *cam, int pixel, float im_w, float im_h,
                float *out);
int launch_proj_bwd(smplb_ctx *c, int B, int N, const float *X, const float *cam, const float *d_out, int pixel,
                    float im_w, float im_h, float gscale, const long long *den, int accumulate_cam, float *d_X,
                    float *d_cam, const int *cnt = nullptr, float denom = 1.0f);
// k_extra.cu
int launch_silhouette_csr(smplb_ctx *c, int B, int H, int W, const float *seg, float *points, int cap, int *offsets,
                          int *counts_scratch);
int launch_silhouette_fill(smplb_ctx *c, int B, int H, int W, const float *seg, float *points, int cap, const int *offsets);
int launch_kcs(smplb_ctx *c, int N, int K, const float *joints, const float *Cm, float *kcs);
int launch_kcs_bwd(smplb_ctx *c, int N, int K, const float *joints, const float *Cm, const float *dK, float *d_joints);
int launch_interp(smplb_ctx *c, size_t total, int row, const float *fake, const float *real, const float *alpha, float *out);
int launch_critic_inputs(smplb_ctx *c, int N, int K, const float *fj, const float *rj, const float *aj, const float *fs,
                         const float *rs, const float *as_, const float *fR, const float *rR, const float *aR,
                         const float *Cm, float *oj, float *okcs, float *os, float *oR);
int launch_critic_gp(smplb_ctx *c, int M, int K, long long M_total, const float *joints, const float *Cm, const float *g_kcs,
                     const float *g_j, const float *g_s, const float *g_R, float *gj_total, float *part,
                     unsigned int *ticket, float *col_sums, float *penalty);
// k_loss.cu
int launch_kp_loss(smplb_ctx *c, int B, int K, const float *kp_gt, const float *kp_pred, float *dkp, float *part,
                   int *cnt);
int launch_reduce_kp(smplb_ctx *c, int B, const float *part, const int *cnt, float *abs_sum, long long *num_present,
                     float *cnt_as_float);
int launch_mesh_loss(smplb_ctx *c, int B, int V, const float *pts, const int *offsets, int P, const float *sil_pred,
                     float *loss, float *d_sil_pred, int *cnt_scratch, float *part_scratch, int *ind_ab, int *ind_ba, bool finish_grad = true);
int launch_finalize_loss(smplb_ctx *c, float w_kp, float w_mesh, long long count_override, int have_mesh,
                         float *loss_parts);
int launch_gp_colsum(smplb_ctx *c, int M, const float *g0, const float *g1, const float *g2, const float *g3,
                     float *col_sums);
int launch_gp_final(smplb_ctx *c, long long M_total, const float *col_sums, float *penalty);
int launch_gp_bwd(smplb_ctx *c, int M, long long M_total, const float *col_sums, float *d0, float *d1, float *d2,
                  float *d3);
// k_exchange.cu
int exchange_preload();
int launch_count_exchange(smplb_ctx *c, int B, const float *kp_gt, long long count_override, int mode, long long *den);
int launch_reduce_exchange_finalize(smplb_ctx *c, int B, const float *part, float w_kp, float w_mesh, int have_mesh,
                                    const long long *den, float *loss_parts);
int launch_finalize_den(smplb_ctx *c, float w_kp, float w_mesh, int have_mesh, const long long *den, float *loss_parts);
