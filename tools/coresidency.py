"""Can a small kernel run while the persistent tensor-core kernel of another context is resident?
ctx0 runs steps (pose_fwd, body_fwd_tc, ...); ctx1 launches stand-alone batch_rodrigues kernels
back to back; the merged timeline shows whether they make progress during body_fwd_tc."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa: E402,F401
from hpe_b200 import synthetic  # noqa: E402
from hpe_b200._lib import check, lib  # noqa: E402
from hpe_b200.tf_smpl.batch_smpl import SMPL  # noqa: E402

B = 4096
model = synthetic.make_model(seed=0)
e0, e1 = SMPL(model, max_batch=B), SMPL(model, max_batch=B)
inp = synthetic.make_inputs(B, seed=1000)
d0 = {k: e0.ctx.to_device(v) for k, v in inp.items()}
th = e1.ctx.to_device(inp["theta"].reshape(-1, 3)[:4096])
Rout = e1.ctx.to_device(np.zeros((4096, 9), np.float32))
out = {}
e0.ctx.debug_set("fused", int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for i in range(3):
    e0.step(d0["beta"], d0["theta"], d0["cam"], d0["kp_gt"], out=out)
    check(lib().smplb_rodrigues(e1.ctx.handle, 4096, th.ptr, Rout.ptr, 1))
e0.ctx.sync(); e1.ctx.sync()
e0.ctx.profile(2); e1.ctx.profile(2)
e0.step(d0["beta"], d0["theta"], d0["cam"], d0["kp_gt"], out=out)
for i in range(12):
    check(lib().smplb_rodrigues(e1.ctx.handle, 4096, th.ptr, Rout.ptr, 1))
rows = [(t0, t1, 0, n) for n, t0, t1 in e0.ctx.profile_trace()] + [(t0, t1, 1, n) for n, t0, t1 in e1.ctx.profile_trace()]
rows.sort()
base = rows[0][0]
for t0, t1, k, n in rows:
    print("%8.1f %8.1f %7.1f  %d %s%s" % ((t0 - base) * 1e3, (t1 - base) * 1e3, (t1 - t0) * 1e3, k, "      " * k, n))
