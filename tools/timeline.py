"""Timeline of the kernels of N contexts running steps back to back (events around every launch,
streams NOT serialised): where the step time goes when the contexts overlap.
Usage: python tools/timeline.py [B] [contexts] [steps] [fused variant]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa: E402,F401
from hpe_b200 import synthetic  # noqa: E402
from hpe_b200.tf_smpl.batch_smpl import SMPL  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NE = int(sys.argv[2]) if len(sys.argv) > 2 else 2
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 4
FUSED = int(sys.argv[4]) if len(sys.argv) > 4 else 1
model = synthetic.make_model(seed=0)
engines = [SMPL(model, max_batch=B) for _ in range(NE)]
for e in engines:
    e.ctx.debug_set("fused", FUSED)
inp = synthetic.make_inputs(B, seed=1000)
dev = [{k: e.ctx.to_device(v) for k, v in inp.items()} for e in engines]
outs = [{} for _ in engines]
for i in range(3 * NE):
    e = i % NE
    engines[e].step(dev[e]["beta"], dev[e]["theta"], dev[e]["cam"], dev[e]["kp_gt"], out=outs[e])
for e in engines:
    e.ctx.sync()
for e in engines:
    e.ctx.profile(2)
for i in range(STEPS * NE):
    e = i % NE
    engines[e].step(dev[e]["beta"], dev[e]["theta"], dev[e]["cam"], dev[e]["kp_gt"], out=outs[e])
rows = []
for k, e in enumerate(engines):
    rows += [(t0, t1, k, name) for name, t0, t1 in e.ctx.profile_trace()]
    e.ctx.profile(0)
rows.sort()
base = rows[0][0]
print("%8s %8s %7s  ctx kernel   (us; start = the event before the launch, which may be queued behind other work)" % ("start", "end", "dur"))
for t0, t1, k, name in rows:
    print("%8.1f %8.1f %7.1f  %d   %s%s" % ((t0 - base) * 1e3, (t1 - base) * 1e3, (t1 - t0) * 1e3, k, "    " * k, name))
span = (rows[-1][1] - base) * 1e3
print("span %.1f us for %d steps -> %.1f us/step, %.2f M meshes/s" % (span, STEPS * NE, span / (STEPS * NE), B * STEPS * NE / span))
