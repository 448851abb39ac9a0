/*
 * smplb.h -- C ABI of libsmplb.so: the batched SMPL body model, weak-perspective
 * projection, reprojection losses and their backward passes as hand-written
 * sm_100a CUDA kernels.
 *
 * The reference (maxpit/human-pose-estimation) has NO plugin / operator / FFI
 * layer for this path: it is plain Python on tf.Tensors.
 * Each entry point below therefore cites the reference *Python* interface it
 * replaces (paths relative to the reference checkout); the ctypes binding a
 * maintainer adds is shown in INTEGRATION.md and shipped in
 * human-pose-estimation_b200/.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SMPLB_E* code on failure;
 *     smplb_last_error() returns the message of the calling thread's last error.
 *     No C++ exceptions cross the ABI.  There is no CPU fallback: creating a
 *     context on anything but an sm_100 device fails.
 *   - all tensors are fp32, row-major, batch-first, exactly as the reference's.
 *   - `mem` says where EVERY data pointer of that call lives:
 *     SMPLB_HOST (pageable or pinned host memory; the call copies in, runs,
 *     copies out and returns when the outputs are valid) or SMPLB_DEVICE
 *     (device memory from smplb_malloc; the call is asynchronous on the
 *     context's stream, order with smplb_sync or smplb_timer_*).  SMPLB_HOST_ASYNC is
 *     SMPLB_HOST without the final synchronisation: buffers must be pinned
 *     (smplb_host_alloc) and outputs are valid after smplb_sync -- lets a caller keep two
 *     contexts in flight so copies of one step overlap the kernels of the other.
 *   - a pointer documented "may be NULL" is an optional output/input.
 *   - a context is bound to one device and is not thread-safe.  Callers order against ONE stream,
 *     the context's main stream; internally a step also uses side streams (the 6890-vertex kernel,
 *     the tcgen05 GEMMs and the loss reduction overlap the per-body kernels), all of which are
 *     joined back into the main stream before the call returns.
 */
#ifndef SMPLB_H_
#define SMPLB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMPLB_VERSION 2

#define SMPLB_HOST 0
#define SMPLB_DEVICE 1
#define SMPLB_HOST_ASYNC 2 /* host buffers (pinned), no synchronisation at return */

#define SMPLB_OK 0
#define SMPLB_EINVAL (-1)   /* bad argument                         */
#define SMPLB_ECUDA (-2)    /* CUDA runtime / driver error          */
#define SMPLB_EDEVICE (-3)  /* no sm_100 device                     */
#define SMPLB_ESTATE (-4)   /* backward without a matching forward  */
#define SMPLB_ENCCL (-5)    /* NCCL unavailable or failed           */

/* flags of smplb_smpl_forward / smplb_step */
#define SMPLB_STEP_KEEP_VERTS 1 /* compute verts even when the verts pointer is NULL and keep them in a
                                   device workspace of the context (smplb_last_verts): what a trainer
                                   wants -- the 339 MB stay in HBM for the renderer / mesh loss and only
                                   the small results cross PCIe */

#define SMPLB_NUM_JOINTS 24
#define SMPLB_NUM_POSE_BASIS 207
#define SMPLB_MAX_KEYPOINTS 32
#define SMPLB_GP_FLOATS 428 /* 13*13 + 14*3 + 10 + 23*9 */

typedef struct smplb_ctx smplb_ctx;

/* Host-side model constants in the layout SMPL.__init__ produces
 * (src/tf_smpl/batch_smpl.py:26-86). */
typedef struct smplb_model {
  int32_t num_verts;            /* V, 6890 for SMPL                                   */
  int32_t num_betas;            /* 10                                (:41)            */
  int32_t num_keypoints;        /* K: 19 cocoplus, 14 lsp            (:75-81)         */
  int32_t reserved;
  const float *v_template;      /* [V,3]                             (:34-38)         */
  const float *shapedirs;       /* [num_betas, 3V], column 3v+c      (:44-47)         */
  const float *posedirs;        /* [207, 3V], column 3v+c            (:57-62)         */
  const float *J_regressor;     /* [V,24]                            (:50-54)         */
  const float *weights;         /* [V,24]                            (:68-72)         */
  const float *joint_regressor; /* [V,K]                             (:75-81)         */
  const int32_t *parents;       /* [24], parents[0] = -1             (:65)            */
} smplb_model;

const char *smplb_last_error(void);
int smplb_version(void);

/* ---- lifecycle: SMPL.__init__ (batch_smpl.py:26-86) ------------------------------- */
int smplb_create(smplb_ctx **ctx, const smplb_model *model, int device, int max_batch);
int smplb_destroy(smplb_ctx *ctx);

/* ---- memory / ordering (so a pure-Python host needs neither torch nor cupy) ------- */
int smplb_malloc(smplb_ctx *ctx, void **dptr, size_t bytes);
int smplb_free(smplb_ctx *ctx, void *dptr);
int smplb_host_alloc(void **hptr, size_t bytes);   /* pinned */
int smplb_host_free(void *hptr);
int smplb_memcpy_h2d(smplb_ctx *ctx, void *dst, const void *src, size_t bytes); /* async on ctx stream */
int smplb_memcpy_d2h(smplb_ctx *ctx, void *dst, const void *src, size_t bytes); /* async on ctx stream */
int smplb_memset(smplb_ctx *ctx, void *dst, int value, size_t bytes);
int smplb_sync(smplb_ctx *ctx);
/* Orders everything enqueued on `ctx` from now on after the work already enqueued on `other`
 * (an event wait between the two contexts' streams; no host synchronisation).        */
int smplb_order_after(smplb_ctx *ctx, smplb_ctx *other);
/* Writes `bytes` of a private scratch buffer (bench: evict L2 between timed steps). */
int smplb_flush_l2(smplb_ctx *ctx, size_t bytes);
/* CUDA-event timers on the context's stream; slot in [0,16). */
int smplb_timer_start(smplb_ctx *ctx, int slot);
int smplb_timer_stop(smplb_ctx *ctx, int slot);
int smplb_timer_elapsed_ms(smplb_ctx *ctx, int slot, float *ms); /* synchronises on the stop event */
/* Kernels launched by this context since creation (bench: gpu_launches). */
int smplb_launch_count(smplb_ctx *ctx, int64_t *count);
/* Per-kernel accumulated device time, measured with events around every launch when
 * enabled (debug/bench breakdown only; serialises nothing but adds event overhead). */
int smplb_profile_enable(smplb_ctx *ctx, int on);   /* 1: per-kernel times (serialised streams); 2: timeline trace */
int smplb_profile_read(smplb_ctx *ctx, char *buf, size_t buflen); /* "name ms count\n" lines (mode 2: "name start_ms end_ms" per launch); resets */
/* (test / tuning hooks live in the private header csrc/smplb_debug.h) */

/* ---- SMPL.__call__(beta, theta, get_skin) (batch_smpl.py:88-160) ------------------ *
 * beta [B,10], theta [B,72] -> verts [B,V,3] (may be NULL == get_skin False),
 * joints [B,K,3], Rs [B,24,3,3] (may be NULL), J_transformed [B,24,3] (may be NULL;
 * the attribute batch_smpl.py:135 sets).  Saves what smplb_smpl_backward needs.
 * flags: 0 or SMPLB_STEP_KEEP_VERTS.                                                  */
int smplb_smpl_forward(smplb_ctx *ctx, int B, const float *beta, const float *theta, float *verts,
                       float *joints, float *Rs, float *J_transformed, int flags, int mem);

/* Backward of the last smplb_smpl_forward on this context (same B).  Replaces TF
 * autodiff through batch_smpl.py:88-160 (src/trainer.py:383,502).  Upstream gradients
 * d_verts [B,V,3], d_joints [B,K,3], d_Rs [B,24,3,3] may each be NULL (== zero).
 * Outputs d_beta [B,10], d_theta [B,72].                                             */
int smplb_smpl_backward(smplb_ctx *ctx, int B, const float *d_verts, const float *d_joints,
                        const float *d_Rs, float *d_beta, float *d_theta, int mem);

/* Device pointer to the verts [B,V,3] of the last forward / step (the caller's buffer in
 * device mode, else a workspace of the context that stays valid until the next call); NULL if
 * the last call did not compute verts.  With SMPLB_STEP_KEEP_VERTS in the call's flags a call
 * that passes verts == NULL still computes them and keeps them on the device.          */
int smplb_last_verts(smplb_ctx *ctx, const float **dptr);

/* ---- src/tf_smpl/batch_lbs.py stand-alone entry points ---------------------------- */
/* batch_rodrigues(theta) (batch_lbs.py:42-64): theta [N,3] -> R [N,3,3].            */
int smplb_rodrigues(smplb_ctx *ctx, int N, const float *theta, float *R, int mem);
/* batch_global_rigid_transformation(Rs, Js, parent) (batch_lbs.py:91-152), parents
 * from the context: Rs [B,24,3,3], Js [B,24,3] -> new_J [B,24,3], A [B,24,4,4].     */
int smplb_global_rigid(smplb_ctx *ctx, int B, const float *Rs, const float *Js, float *new_J, float *A,
                       int mem);
/* batch_skew(vec) (batch_lbs.py:15-39): vec [N,3] -> [N,3,3].                        */
int smplb_skew(smplb_ctx *ctx, int N, const float *vec, float *out, int mem);
/* batch_lrotmin(theta) (batch_lbs.py:67-88, unused by the reference): theta [B,72] ->
 * (Rodrigues(theta[:,3:]) - I) as [B,207].                                           */
int smplb_lrotmin(smplb_ctx *ctx, int B, const float *theta, float *out, int mem);

/* ---- src/tf_smpl/projection.py ---------------------------------------------------- */
/* batch_orth_proj_idrot(X, camera) (projection.py:23-33): X [B,N,3], cam [B,3] ->
 * out [B,N,2] = s * (X_xy + t).                                                      */
int smplb_orth_proj(smplb_ctx *ctx, int B, int N, const float *X, const float *cam, float *out, int mem);
/* reproject_vertices(verts, cam, im_size) (projection.py:45-56): -> pixels [B,N,2]. */
int smplb_reproject_vertices(smplb_ctx *ctx, int B, int N, const float *verts, const float *cam,
                             float im_w, float im_h, float *out, int mem);
/* Backward of either projection: d_out [B,N,2] -> d_X [B,N,3] (may be NULL; z = 0),
 * d_cam [B,3] (may be NULL).  pixel != 0 selects reproject_vertices.                 */
int smplb_proj_backward(smplb_ctx *ctx, int B, int N, const float *X, const float *cam, const float *d_out,
                        int pixel, float im_w, float im_h, float *d_X, float *d_cam, int mem);

/* ---- src/ops.py ------------------------------------------------------------------- */
/* kp_reprojection_loss(kp_gt, kp_pred) (ops.py:35-47).  Returns the NUMERATOR
 * sum(vis * |gt - pred|) and the integer count 2 * #{vis != 0} separately so that
 * batch shards can be all-reduced exactly; loss = abs_sum / num_present (0 if 0).
 * kp_gt [B,K,3], kp_pred [B,K,2]; d_kp_pred [B,K,2] may be NULL, else receives
 * vis * sign(pred - gt) (UNSCALED; multiply by 1 / global num_present).             */
int smplb_kp_loss(smplb_ctx *ctx, int B, int K, const float *kp_gt, const float *kp_pred, float *abs_sum,
                  int64_t *num_present, float *d_kp_pred, int mem);

/* mesh_reprojection_loss (ops.py:117-137) with find_nearest_neighbors /
 * bidirectional_dist (ops.py:60-102).  points_xy [P,2] are the (x,y) = (col,row)
 * silhouette pixels of all images concatenated, offsets [B+1] (int32) delimits each
 * image's rows (the CSR form of the reference's [P,3] (n,row,col) list), sil_pred
 * [B,V,2].  loss gets sum_i bidirectional_dist_i / (3 + V).  d_sil_pred [B,V,2] may be
 * NULL.  An image with no pixels contributes 0.  ind_ab [P] / ind_ba [B,V] (may be NULL)
 * receive find_nearest_neighbors' indices (ops.py:68-69; ind_ab is local to the image's
 * vertex list, ind_ba local to the image's pixel list, -1 for an empty image).       */
int smplb_mesh_reproj_loss(smplb_ctx *ctx, int B, int V, const float *points_xy, const int32_t *offsets,
                           int P, const float *sil_pred, float *loss, float *d_sil_pred, int32_t *ind_ab,
                           int32_t *ind_ba, int mem);

/* compute_gradient_penalty(gradients) (ops.py:153-172): g0 [M,13,13], g1 [M,14,3],
 * g2 [M,10], g3 [M,23,3,3].  col_sums [428] (may be NULL) receives sum over M of every
 * column -- the vector a multi-GPU caller all-reduces before the norm; penalty gets
 * sum_i (1 - ||col_sums_i / M||)^2.                                                  */
int smplb_gradient_penalty(smplb_ctx *ctx, int M, const float *g0, const float *g1, const float *g2,
                           const float *g3, float *penalty, float *col_sums, int mem);
/* Same reduction from already-summed columns (after an all-reduce); M_total rows.   */
int smplb_gradient_penalty_from_sums(smplb_ctx *ctx, int64_t M_total, const float *col_sums, float *penalty,
                                     int mem);
/* d penalty / d g_i, broadcast over the M rows; any d_g may be NULL.                */
int smplb_gradient_penalty_backward(smplb_ctx *ctx, int M, int64_t M_total, const float *col_sums, float *d_g0,
                                    float *d_g1, float *d_g2, float *d_g3, int mem);

/* ---- neighbours of the path ------------------------------------------- *
 * tf.cast(tf.where(seg > 0)[:, :3], float32) (src/trainer.py:291,443) as an on-device ordered
 * compaction into the CSR form smplb_mesh_reproj_loss / smplb_step take: seg [B,H,W] (NHWC with
 * C = 1), points_xy [cap,2] receives (x = col, y = row) in row-major order per image, offsets
 * [B+1]; points beyond cap are dropped (offsets still report the true counts).          */
int smplb_silhouette_csr(smplb_ctx *ctx, int B, int H, int W, const float *seg, float *points_xy, int cap,
                         int32_t *offsets, int mem);
/* get_kcs(joints, C_matrix) (src/models.py:123-139): joints [N,K,3] (first 14 used), C [14,13]
 * -> kcs [N,13,13]; and its backward d_kcs [N,13,13] -> d_joints [N,K,3].             */
int smplb_kcs(smplb_ctx *ctx, int N, int K, const float *joints, const float *C, float *kcs, int mem);
int smplb_kcs_backward(smplb_ctx *ctx, int N, int K, const float *joints, const float *C, const float *d_kcs,
                       float *d_joints, int mem);
/* Critic-input interpolation fake + alpha * (real - fake), alpha [N] per row of length `row`
 * (src/trainer.py:551-557).                                                           */
int smplb_interpolate(smplb_ctx *ctx, int N, int row, const float *fake, const float *real, const float *alpha,
                      float *out, int mem);

/* The critic's interpolated inputs in one launch (src/trainer.py:548-557): x_hat = fake + alpha * (real - fake) with
 * ELEMENT-wise alpha (tf.random.uniform(x.shape)) for the 3-D joints [N,K,3] (K >= 14), the shapes [N,10] and the
 * rotations [N,23,3,3], plus kcs [N,13,13] = get_kcs(x_hat joints, C) (src/models.py:123-139).                   */
int smplb_critic_inputs(smplb_ctx *ctx, int N, int K, const float *fake_joints, const float *real_joints,
                        const float *alpha_joints, const float *fake_shapes, const float *real_shapes,
                        const float *alpha_shapes, const float *fake_Rs, const float *real_Rs, const float *alpha_Rs,
                        const float *C, float *joints, float *kcs, float *shapes, float *Rs, int mem);
/* tf.gradients(out, [kcs, joints, shapes, Rs]) + compute_gradient_penalty (src/trainer.py:566-572, src/ops.py:153-172)
 * in one launch, given the critic's PARTIAL derivatives g_kcs [M,13,13], g_joints [M,14,3] (direct path only),
 * g_shapes [M,10], g_Rs [M,23,3,3] and the interpolated joints [M,K,3]: the joints gradient gains the path through
 * get_kcs (its backward, per row), col_sums [428] (may be NULL) and penalty as smplb_gradient_penalty;
 * g_joints_total [M,14,3] (may be NULL) receives the total joints gradient.  M_total: rows over all ranks.      */
int smplb_critic_gradient_penalty(smplb_ctx *ctx, int M, int K, int64_t M_total, const float *joints, const float *C,
                                  const float *g_kcs, const float *g_joints, const float *g_shapes, const float *g_Rs,
                                  float *penalty, float *col_sums, float *g_joints_total, int mem);

/* ---- the benchmarked fused call: what one generator stage of Trainer.train_step does
 * with SMPL, projection and losses (src/trainer.py:404-450) plus its backward
 * (src/trainer.py:502).
 *   in : beta [B,10], theta [B,72], cam [B,3], kp_gt [B,K,3]
 *        optional silhouettes: points_xy [P,2], offsets [B+1] (NULL == no mesh loss)
 *        w_kp, w_mesh: loss weights (src/config.py:67-68: 60 and 0.001)
 *        img_size: side of the square image the silhouettes live in (src/config.py:36: 224)
 *        kp_count_override: if > 0 use it as the GLOBAL num_present (multi-GPU: the
 *        all-reduced count, which depends only on kp_gt) else the local count.
 *   out: verts [B,V,3] (may be NULL only if no mesh loss), joints [B,K,3], Rs [B,24,3,3]
 *        (may be NULL), kp_pred [B,K,2] (may be NULL),
 *        loss_parts [4] = {kp abs_sum, kp num_present, mesh loss sum, total weighted loss};
 *        with the batch sharded over ranks (smplb_comm_p2p_attach* or smplb_comm_init) the
 *        first three are summed over the ranks inside the call -- the visibility count at
 *        the start of the step, the numerators next to the backward -- and the gradients use
 *        the global count
 *        d_beta [B,10], d_theta [B,72], d_cam [B,3] (all three may be NULL == forward only)
 *   flags: 0 or SMPLB_STEP_KEEP_VERTS                                                      */
int smplb_step(smplb_ctx *ctx, int B, const float *beta, const float *theta, const float *cam,
               const float *kp_gt, const float *points_xy, const int32_t *offsets, int P, float w_kp,
               float w_mesh, float img_size, int64_t kp_count_override, float *verts, float *joints, float *Rs,
               float *kp_pred, float *loss_parts, float *d_beta, float *d_theta, float *d_cam, int flags,
               int mem);

/* The same step with the silhouettes given as the dense mask the trainer holds, seg [B,H,W] (NHWC with C = 1;
 * src/trainer.py:443 builds the point list with tf.where(seg > 0) before every mesh_reprojection_loss): the
 * compaction runs on the device inside the call, nothing passes through the host but one 4-byte count.       */
int smplb_step_seg(smplb_ctx *ctx, int B, const float *beta, const float *theta, const float *cam,
                   const float *kp_gt, const float *seg, int H, int W, float w_kp, float w_mesh, float img_size,
                   int64_t kp_count_override, float *verts, float *joints, float *Rs, float *kp_pred,
                   float *loss_parts, float *d_beta, float *d_theta, float *d_cam, int flags, int mem);

/* ---- multi-GPU: one process per GPU; the batch shards over the ranks and the only exchange is
 * the sum of the visibility count and of the loss numerators (the
 * reference is single-device, src/trainer.py:352).  Two transports:
 *  (1) mailboxes in peer GPU memory, written / polled by the loss kernels themselves over
 *      NVLink (no collective launch): every rank exports the handle of its context's mailbox,
 *      the host exchanges the 64-byte handles by any means (torch.distributed, MPI, a file) and
 *      attaches them.  All ranks must issue the steps of their contexts in the same order.   */
int smplb_comm_p2p_export(smplb_ctx *ctx, void *handle64);            /* CUDA IPC handle, 64 bytes */
int smplb_comm_p2p_attach(smplb_ctx *ctx, int nranks, int rank, const void *handles /* nranks x 64 B */);
/* the same when the ranks are contexts of ONE process (peers[rank] == ctx)              */
int smplb_comm_p2p_attach_local(smplb_ctx *ctx, int nranks, int rank, smplb_ctx *const *peers);
/* 0, or 1 once a rank waited longer than the timeout for a peer (that step's loss is NaN) */
int smplb_comm_status(smplb_ctx *ctx, int *status);
/* (2) NCCL (dlopen'ed, not a link-time dependency): used when no mailboxes are attached.   */
int smplb_comm_unique_id(void *id128);                 /* 128-byte NCCL unique id (rank 0)   */
int smplb_comm_init(smplb_ctx *ctx, int nranks, int rank, const void *id128);
int smplb_comm_allreduce_sum(smplb_ctx *ctx, float *dev_buf, int count); /* in place, ctx stream */
int smplb_comm_destroy(smplb_ctx *ctx);               /* detaches both transports            */

#ifdef __cplusplus
}
#endif
#endif /* SMPLB_H_ */
