"""BASELINE config 5: inference path (src/predictor.py shape) -- SMPL forward only,
batch sweep for latency / throughput.  Device-resident I/O, CUDA events."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL

maxB = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model = synthetic.make_model(seed=0)
smpl = SMPL(model, max_batch=maxB)
ctx = smpl.ctx
rows = []
B = 1
while B <= maxB:
    inp = synthetic.make_inputs(B, seed=7)
    db, dt = ctx.to_device(inp["beta"]), ctx.to_device(inp["theta"])
    for skin in (True, False):
        for _ in range(3):
            r = smpl(db, dt, get_skin=skin)
        ctx.sync()
        n = max(3, min(50, 200000 // max(B, 1)))
        ctx.timer_start(0)
        for _ in range(n):
            r = smpl(db, dt, get_skin=skin)
        ctx.timer_stop(0)
        ms = ctx.timer_ms(0) / n
        rows.append({"B": B, "get_skin": skin, "ms": ms, "meshes_per_s": B / ms * 1e3,
                     "gbs": (B * 84400 + 19870760) / ms / 1e6 if skin else None})
        print(json.dumps(rows[-1]))
        del r
    del db, dt
    B *= 4
