import os, sys
sys.path.insert(0, "/root/repo")
import hpe_b200
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
B = 4096
smpl = SMPL(synthetic.make_model(seed=0), max_batch=B)
inp = synthetic.make_inputs(B, seed=1000)
d = {k: smpl.ctx.to_device(v) for k, v in inp.items()}
out = {}
for it in range(4):
    smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], want_verts=False, out=out)
smpl.ctx.sync()
