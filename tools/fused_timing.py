"""Run two steps at B = 4096 (for builds of k_body_pair.cu / k_body_tc.cu with -DFB_TIMING, which print per-warp
wait/busy cycle counts of CTAs 0 and 77).  Usage: python tools/fused_timing.py [fused variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa: E402,F401
from hpe_b200 import synthetic  # noqa: E402
from hpe_b200.tf_smpl.batch_smpl import SMPL  # noqa: E402

B = 4096
smpl = SMPL(synthetic.make_model(seed=0), max_batch=B)
ctx = smpl.ctx
inp = synthetic.make_inputs(B, seed=1000)
d = {k: ctx.to_device(v) for k, v in inp.items()}
out = {}
ctx.debug_set("fused", int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for it in range(2):
    smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], out=out)
ctx.sync()
