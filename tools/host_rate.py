"""Is the step loop limited by the host?  Times the enqueue loop (no sync inside) and the total
including the final sync, for N contexts in flight.  Usage: python tools/host_rate.py [B] [contexts] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa: E402,F401
from hpe_b200 import synthetic  # noqa: E402
from hpe_b200.tf_smpl.batch_smpl import SMPL  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NE = int(sys.argv[2]) if len(sys.argv) > 2 else 3
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
model = synthetic.make_model(seed=0)
engines = [SMPL(model, max_batch=B) for _ in range(NE)]
inp = synthetic.make_inputs(B, seed=1000)
dev = [{k: e.ctx.to_device(v) for k, v in inp.items()} for e in engines]
outs = [{} for _ in engines]


def run(n):
    for i in range(n):
        e = i % NE
        engines[e].step(dev[e]["beta"], dev[e]["theta"], dev[e]["cam"], dev[e]["kp_gt"], out=outs[e])


run(30)
for e in engines:
    e.ctx.sync()
t0 = time.perf_counter()
run(STEPS)
t1 = time.perf_counter()
for e in engines:
    e.ctx.sync()
t2 = time.perf_counter()
print("B=%d, %d contexts: enqueue loop %.1f us/step, with final sync %.1f us/step" % (B, NE, (t1 - t0) / STEPS * 1e6, (t2 - t0) / STEPS * 1e6))
