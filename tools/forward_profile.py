import sys
sys.path.insert(0, "/root/repo")
import hpe_b200
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
smpl = SMPL(synthetic.make_model(seed=0), max_batch=4096)
ctx = smpl.ctx
for B in (64, 128, 192, 256, 512):
    inp = synthetic.make_inputs(B, seed=1000)
    db, dt = ctx.to_device(inp["beta"]), ctx.to_device(inp["theta"])
    v, j, r = ctx.empty((B, 6890, 3)), ctx.empty((B, 19, 3)), ctx.empty((B, 24, 3, 3))
    for i in range(5): smpl.forward_into(db, dt, B, v, j, r)
    ctx.sync()
    ctx.timer_start(0)
    for i in range(200): smpl.forward_into(db, dt, B, v, j, r)
    ctx.timer_stop(0)
    t = ctx.timer_ms(0) / 200
    ctx.profile(True)
    for i in range(10): smpl.forward_into(db, dt, B, v, j, r)
    prof = ctx.profile_read(); ctx.profile(False)
    print(B, "%.1f us" % (t * 1e3), {k: round(ms / n * 1e3, 1) for k, (ms, n) in prof.items()})
