"""Timeline of the keypoint chain alone (no verts) with 3 contexts in flight."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
B, NE, STEPS = 4096, 3, 3
model = synthetic.make_model(seed=0)
eng = [SMPL(model, max_batch=B) for _ in range(NE)]
inp = synthetic.make_inputs(B, seed=1000)
dev = [{k: e.ctx.to_device(v) for k, v in inp.items()} for e in eng]
outs = [{} for _ in eng]
step = lambda k: eng[k].step(dev[k]["beta"], dev[k]["theta"], dev[k]["cam"], dev[k]["kp_gt"], want_verts=False, out=outs[k])
for i in range(3 * NE): step(i % NE)
for e in eng: e.ctx.sync()
for e in eng: e.ctx.profile(2)
for i in range(STEPS * NE): step(i % NE)
rows = []
for k, e in enumerate(eng):
    rows += [(t0, t1, k, name) for name, t0, t1 in e.ctx.profile_trace()]
    e.ctx.profile(0)
rows.sort()
base = rows[0][0]
for t0, t1, k, name in rows:
    print("%8.1f %8.1f %7.1f  %d   %s%s" % ((t0 - base) * 1e3, (t1 - base) * 1e3, (t1 - t0) * 1e3, k, "    " * k, name))
span = (rows[-1][1] - base) * 1e3
print("span %.1f us for %d steps -> %.1f us/step" % (span, STEPS * NE, span / (STEPS * NE)))
