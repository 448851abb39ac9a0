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
#ifdef __cplusplus
}
#endif
#endif
