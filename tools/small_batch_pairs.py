import sys
sys.path.insert(0, "/root/repo")
import hpe_b200
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
smpl = SMPL(synthetic.make_model(seed=0), max_batch=4096)
ctx = smpl.ctx
for B in (256, 384, 512, 768, 1024, 2048):
    inp = synthetic.make_inputs(B, seed=1000)
    db, dt = ctx.to_device(inp["beta"]), ctx.to_device(inp["theta"])
    v, j, r = ctx.empty((B, 6890, 3)), ctx.empty((B, 19, 3)), ctx.empty((B, 24, 3, 3))
    res = []
    for pairs in (0, 68, 64, 60, 54, 48):
        ctx.debug_set("body_pairs", pairs)
        for i in range(5): smpl.forward_into(db, dt, B, v, j, r)
        ctx.sync()
        ctx.timer_start(0)
        for i in range(300): smpl.forward_into(db, dt, B, v, j, r)
        ctx.timer_stop(0)
        res.append("%d: %.1f" % (pairs, ctx.timer_ms(0) / 300 * 1e3))
    print("B=%d forward call us by vertex-kernel pairs (0 = all 74): %s" % (B, ", ".join(res)))
ctx.debug_set("body_pairs", 0)
