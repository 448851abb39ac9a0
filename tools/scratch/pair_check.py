import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hpe_b200
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
smpl = SMPL(synthetic.make_model(seed=0), max_batch=512)
ctx = smpl.ctx
for B in (1, 7, 96, 203, 500):
    inp = synthetic.make_inputs(B, seed=500 + B)
    ctx.debug_set("fused", 1)
    v0, j0, _ = smpl(inp["beta"], inp["theta"], get_skin=True)
    ctx.debug_set("fused", 7)
    v1, j1, _ = smpl(inp["beta"], inp["theta"], get_skin=True)
    err = np.abs(v1 - v0).max() / np.abs(v0).max()
    print("B", B, "rel err", err, "finite", np.isfinite(v1).all(), flush=True)
