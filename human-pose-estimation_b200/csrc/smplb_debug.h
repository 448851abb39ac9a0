/* Private test / tuning hooks of libsmplb.so.  NOT part of the public ABI (include/smplb.h): the
 * symbols are exported for the repo's own tests and tools and may change without notice. */
#ifndef SMPLB_DEBUG_H_
#define SMPLB_DEBUG_H_
#include "smplb.h"
#ifdef __cplusplus
extern "C" {
#endif
/* key = value on one context.  Cross-check paths: "blend_tc" = 0 routes the blend contraction
 * through the FP32 CUDA-core GEMM instead of tcgen05 (default 1); "skin_tc", "fused", "fold",
 * "fold_step", "fold_warp", "compact_bwd", "mesh_grid" likewise.  Scheduling: "overlap", "prio",
 * "body_pairs", "l2_chunk".  Exchange: "comm_backend" (1 = NCCL even with mailboxes attached),
 * "comm_timeout_ms". */
int smplb_debug_set(smplb_ctx *ctx, const char *key, int value);
/* Writes rank `from_rank`'s mailbox entry of exchange `epoch` (1 = the first sharded step after attach; kind 0 = the
 * visibility count in `cnt`, kind 1 = {kp numerator, mesh sum} in v0 / v1) into this context's mailbox, as that
 * rank's push would: single-GPU tests run the ranks one after the other with it. */
int smplb_debug_p2p_inject(smplb_ctx *ctx, int kind, unsigned epoch, int from_rank, float v0, float v1, long long cnt);
#ifdef __cplusplus
}
#endif
#endif
