import sys
sys.path.insert(0, "/root/repo")
import hpe_b200
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
model = synthetic.make_model(seed=0)
smpl = SMPL(model, max_batch=16384)
ctx = smpl.ctx
for B in (4096, 16384):
    inp = synthetic.make_inputs(B, seed=1000)
    db, dt = ctx.to_device(inp["beta"]), ctx.to_device(inp["theta"])
    v, j, r = ctx.empty((B, 6890, 3)), ctx.empty((B, 19, 3)), ctx.empty((B, 24, 3, 3))
    res = []
    for pairs in (0, 70, 68, 66, 64, 60):
        ctx.debug_set("body_pairs", pairs)
        for i in range(5): smpl.forward_into(db, dt, B, v, j, r)
        ctx.sync()
        ctx.timer_start(0)
        n = 100 if B == 4096 else 30
        for i in range(n): smpl.forward_into(db, dt, B, v, j, r)
        ctx.timer_stop(0)
        res.append("%d: %.1f" % (pairs, ctx.timer_ms(0) / n * 1e3))
    print("B=%d forward call us by pairs: %s" % (B, ", ".join(res)), flush=True)
del smpl
# training step, 1 and 3 contexts in flight
B = 4096
for NE in (1, 3):
    eng = [SMPL(model, max_batch=B) for _ in range(NE)]
    inp = synthetic.make_inputs(B, seed=1000)
    dev = [{k: e.ctx.to_device(v) for k, v in inp.items()} for e in eng]
    outs = [{} for _ in eng]
    res = []
    for pairs in (0, 70, 68, 66, 64, 60, 0, 64):
        for e in eng: e.ctx.debug_set("body_pairs", pairs)
        fn = lambda k: eng[k].step(dev[k]["beta"], dev[k]["theta"], dev[k]["cam"], dev[k]["kp_gt"], out=outs[k])
        for i in range(3 * NE): fn(i % NE)
        for e in eng: e.ctx.sync()
        c = eng[0].ctx
        c.timer_start(0)
        for e in eng[1:]: e.ctx.order_after(c)
        N = 600
        for i in range(N): fn(i % NE)
        for e in eng[1:]: c.order_after(e.ctx)
        c.timer_stop(0)
        res.append("%d: %.1f" % (pairs, c.timer_ms(0) / N * 1e3))
    print("step B=4096, %d context(s) in flight, us/step by pairs: %s" % (NE, ", ".join(res)), flush=True)
    del eng
