"""One config-3-like mesh step (B = 1024) for profiling runs (ncu -k regex:k_mesh)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa
from hpe_b200 import ops, synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
smpl = SMPL(synthetic.make_model(seed=0), max_batch=B)
ctx = smpl.ctx
inp = synthetic.make_inputs(B, seed=1000)
d = {k: ctx.to_device(v) for k, v in inp.items()}
seg = synthetic.make_silhouettes(B, seed=2000)
pts, offs = ops.silhouette_csr(synthetic.silhouette_points(seg), B)
sil = (ctx.to_device(pts), ctx.to_device(offs, np.int32))
out = {}
for it in range(3):
    smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], silhouette=sil, out=out)
ctx.sync()
