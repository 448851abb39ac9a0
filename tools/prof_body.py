"""Five forwards (verts + joints) at B = 4096 on one context: the command line ncu profiles the vertex kernel on.
Usage: python tools/prof_body.py [fused variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa: E402,F401
from hpe_b200 import synthetic  # noqa: E402
from hpe_b200.tf_smpl.batch_smpl import SMPL  # noqa: E402

B = 4096
smpl = SMPL(synthetic.make_model(seed=0), max_batch=B)
ctx = smpl.ctx
if len(sys.argv) > 1:
    ctx.debug_set("fused", int(sys.argv[1]))
inp = synthetic.make_inputs(B, seed=1000)
db, dt = ctx.to_device(inp["beta"]), ctx.to_device(inp["theta"])
v, j, r = ctx.empty((B, 6890, 3)), ctx.empty((B, 19, 3)), ctx.empty((B, 24, 3, 3))
for i in range(5):
    smpl.forward_into(db, dt, B, v, j, r)
ctx.sync()
print("ok")
