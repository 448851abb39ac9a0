"""Config-3 step (B = 1024, keypoint + mesh loss + backward) per-kernel breakdown.  Usage: python tools/config3.py [key=value ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa: E402,F401
from hpe_b200 import ops, synthetic  # noqa: E402
from hpe_b200.tf_smpl.batch_smpl import SMPL  # noqa: E402

B = 1024
smpl = SMPL(synthetic.make_model(seed=0), max_batch=B)
ctx = smpl.ctx
for kv in sys.argv[1:]:
    ctx.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
inp = synthetic.make_inputs(B, seed=3000)
seg = synthetic.make_silhouettes(B, seed=3001)
pts, offs = ops.silhouette_csr_device(ctx.to_device(seg.reshape(B, 224, 224)), cap=int(seg.sum()) + 16)
d = {k: ctx.to_device(v) for k, v in inp.items()}
out = {}
for i in range(3):
    smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], silhouette=(pts, offs), out=out)
ctx.sync()
ctx.timer_start(0)
N = 20
for i in range(N):
    smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], silhouette=(pts, offs), out=out)
ctx.timer_stop(0)
print("step: %.3f ms  (%.0f meshes/s)" % (ctx.timer_ms(0) / N, B * N / ctx.timer_ms(0) * 1e3))
ctx.profile(True)
for i in range(5):
    smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], silhouette=(pts, offs), out=out)
for k, (ms, n) in ctx.profile_read().items():
    print("  %-34s %8.1f us" % (k, ms / 5 * 1e3))
print("loss_parts", out["loss_parts"].numpy())
