"""Per-kernel device-time breakdown of one fused step (events around every
launch; smplb_profile_*).  Usage: python tools/breakdown.py [B] [mesh]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa: E402,F401
from hpe_b200 import ops, synthetic  # noqa: E402
from hpe_b200.tf_smpl.batch_smpl import SMPL  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mesh = len(sys.argv) > 2 and sys.argv[2] == "mesh"
model = synthetic.make_model(seed=0)
smpl = SMPL(model, max_batch=B)
ctx = smpl.ctx
inp = synthetic.make_inputs(B, seed=1000)
d = {k: ctx.to_device(v) for k, v in inp.items()}
sil = None
if mesh:
    seg = synthetic.make_silhouettes(B, seed=2000)
    pts, offs = ops.silhouette_csr(synthetic.silhouette_points(seg), B)
    print("silhouette points:", len(pts), "mean per image", len(pts) / B)
    sil = (ctx.to_device(pts), ctx.to_device(offs, np.int32))
if os.environ.get("L2_CHUNK"):
    ctx.debug_set("l2_chunk", int(os.environ["L2_CHUNK"]))
out = {}
for it in range(3):
    smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], silhouette=sil, out=out)
ctx.sync()
ctx.profile(True)
N = 5
for it in range(N):
    smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], silhouette=sil, out=out)
prof = ctx.profile_read()
ctx.profile(False)
tot = sum(v[0] for v in prof.values()) / N
print("B=%d mesh=%s  sum of kernel times per step: %.3f ms -> %.3f M meshes/s" % (B, mesh, tot, B / tot / 1e3))
for k, (ms, n) in prof.items():
    print("  %-28s %9.3f ms/step  (%d launches/step)  %5.1f%%" % (k, ms / N, n // N, 100 * ms / N / tot))
ctx.timer_start(0)
for it in range(N):
    smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], silhouette=sil, out=out)
ctx.timer_stop(0)
ms = ctx.timer_ms(0) / N
print("whole step (events, no per-kernel profiling): %.3f ms -> %.3f M meshes/s" % (ms, B / ms / 1e3))
print("loss_parts", out["loss_parts"].numpy())
