"""Small end-to-end run of every entry point for compute-sanitizer (memcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200
from hpe_b200 import ops, synthetic
from hpe_b200.tf_smpl import batch_lbs, projection
from hpe_b200.tf_smpl.batch_smpl import SMPL

model = synthetic.make_model(seed=0)
smpl = SMPL(model, max_batch=40)
for B in (1, 37):
    inp = synthetic.make_inputs(B, seed=B)
    seg = synthetic.make_silhouettes(B, seed=3, a_range=(4, 7), b_range=(6, 10))
    sil = ops.silhouette_csr(synthetic.silhouette_points(seg), B)
    out = smpl.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
    out = smpl.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], silhouette=sil)
    v, j, R = smpl(inp["beta"], inp["theta"], get_skin=True)
    jj = smpl(inp["beta"], inp["theta"])
    db, dt = smpl.backward(d_verts=np.ones_like(v), d_joints=np.ones_like(j), d_Rs=np.ones_like(R))
    db, dt = smpl.backward(d_joints=np.ones_like(j))
    for key in ("fold", "fold_step", "fused", "blend_tc", "skin_tc", "compact_bwd"):
        smpl.ctx.debug_set(key, 0)
        out = smpl.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
        smpl.ctx.debug_set(key, 1)
    for variant in (2, 3, 5, 6):                      # tuning variants of the fused blend + skinning kernel
        smpl.ctx.debug_set("fused", variant)
        v2, _, _ = smpl(inp["beta"], inp["theta"], get_skin=True)
    smpl.ctx.debug_set("fused", 1)
    batch_lbs.batch_rodrigues(inp["theta"].reshape(-1, 3)); batch_lbs.batch_lrotmin(inp["theta"])
    projection.reproject_vertices(v, inp["cam"], [224., 224.])
    ops.compute_gradient_penalty(synthetic.make_gp_inputs(3 * B))
print("sanitize smoke done")
