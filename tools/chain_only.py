"""Throughput of the keypoint chain alone (no 6890-vertex kernel: verts not requested) and of the vertex kernel alone,
three contexts in flight: what the step would cost if the two overlapped perfectly / not at all."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL
B, NE, N = 4096, 3, 600
model = synthetic.make_model(seed=0)
eng = [SMPL(model, max_batch=B) for _ in range(NE)]
inp = synthetic.make_inputs(B, seed=1000)
dev = [{k: e.ctx.to_device(v) for k, v in inp.items()} for e in eng]
outs = [{} for _ in eng]
vb = [e.ctx.empty((B, 6890, 3)) for e in eng]
jb = [e.ctx.empty((B, 19, 3)) for e in eng]
rb = [e.ctx.empty((B, 24, 3, 3)) for e in eng]
def run(fn, label):
    for i in range(3 * NE): fn(i % NE)
    for e in eng: e.ctx.sync()
    c = eng[0].ctx
    c.timer_start(0)
    for e in eng[1:]: e.ctx.order_after(c)
    for i in range(N): fn(i % NE)
    for e in eng[1:]: c.order_after(e.ctx)
    c.timer_stop(0)
    print("%-44s %.1f us/step" % (label, c.timer_ms(0) / N * 1e3), flush=True)
run(lambda k: eng[k].step(dev[k]["beta"], dev[k]["theta"], dev[k]["cam"], dev[k]["kp_gt"], out=outs[k]), "full step (verts + keypoint chain)")
o2 = [{} for _ in eng]
run(lambda k: eng[k].step(dev[k]["beta"], dev[k]["theta"], dev[k]["cam"], dev[k]["kp_gt"], want_verts=False, out=o2[k]), "keypoint chain only (no verts)")
run(lambda k: eng[k].forward_into(dev[k]["beta"], dev[k]["theta"], B, vb[k], jb[k], rb[k]), "forward with verts (pose + vertex kernel + joints)")
