"""tcgen05 blend GEMM vs the FP32 CUDA-core GEMM and the fp64 oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
from oracle import smpl_numpy as onp

model = synthetic.make_model(seed=0)
smpl = SMPL(model, max_batch=512)
o = onp.SMPL(model, dtype=np.float64)
for B in (8, 129, 300):
    inp = synthetic.make_inputs(B, seed=B)
    smpl.ctx.debug_set("blend_tc", 0)
    v0, j0, _ = smpl(inp["beta"], inp["theta"], get_skin=True)
    smpl.ctx.debug_set("blend_tc", 1)
    v1, j1, _ = smpl(inp["beta"], inp["theta"], get_skin=True)
    n = min(B, 16)
    vr, jr, _ = o(inp["beta"][:n].astype(np.float64), inp["theta"][:n].astype(np.float64), get_skin=True)
    sc = np.abs(vr).max()
    print("B=%d  tc vs sgemm: %.3e   sgemm vs oracle: %.3e   tc vs oracle: %.3e  (scale-relative), finite=%s" % (
        B, np.abs(v1 - v0).max() / sc, np.abs(v0[:n] - vr).max() / sc, np.abs(v1[:n] - vr).max() / sc, np.isfinite(v1).all()))
    bad = np.argwhere(np.abs(v1 - v0) > 1e-3 * sc)
    if len(bad):
        print("   mismatches:", len(bad), "first", bad[:5].tolist(), "rows", np.unique(bad[:, 0])[:10], "verts", np.unique(bad[:, 1])[:10])
