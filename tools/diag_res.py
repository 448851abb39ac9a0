"""Diagnose the resident-tile kernel: forward alone, step without overlap, step with overlap."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
smpl = SMPL(synthetic.make_model(seed=0), max_batch=B)
ctx = smpl.ctx
inp = synthetic.make_inputs(B, seed=1000)
d = {k: ctx.to_device(v) for k, v in inp.items()}
if mode == "fwd":
    for i in range(3):
        v, j, R = smpl(d["beta"], d["theta"], get_skin=True)
        ctx.sync()
        print("fwd", i, "ok", flush=True)
    ctx.debug_set("fused", 7)
    v7, _, _ = smpl(d["beta"], d["theta"], get_skin=True)
    a, b = v.numpy(), v7.numpy()
    print("max diff vs pair kernel", np.abs(a - b).max(), flush=True)
else:
    if mode == "step0":
        ctx.debug_set("overlap", 0)
    out = {}
    for i in range(5):
        smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], out=out)
        ctx.sync()
        print(mode, i, "ok", flush=True)
