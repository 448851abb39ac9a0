"""Clock and board power while ONE kernel mix runs back to back for a few seconds (nvidia-smi, 100 ms samples):
the fused vertex kernel (forward with verts), the keypoint chain alone (step without verts), and -- with SMPLB_LIB
pointing at an ablation build (tools/build_variant.sh) -- the vertex kernel without its MMAs or without its stores.
Usage: python tools/power_probe.py [seconds]"""
import os, subprocess, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa
from hpe_b200 import synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL

SEC = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
B = 4096
smpl = SMPL(synthetic.make_model(seed=0), max_batch=B)
ctx = smpl.ctx
inp = synthetic.make_inputs(B, seed=1000)
d = {k: ctx.to_device(v) for k, v in inp.items()}
v, j, r = ctx.empty((B, 6890, 3)), ctx.empty((B, 19, 3)), ctx.empty((B, 24, 3, 3))
out = {}
rows = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
threading.Thread(target=lambda: [rows.append(l.split(",")) for l in proc.stdout], daemon=True).start()
while not rows:
    time.sleep(0.05)

def probe(label, fn):
    for _ in range(20):
        fn()
    ctx.sync()
    r0 = len(rows)
    ctx.profile(True)
    t0 = time.time()
    n = 0
    while time.time() - t0 < SEC:
        for _ in range(50):
            fn()
        ctx.sync()
        n += 50
    prof = ctx.profile_read()
    ctx.profile(False)
    rs = rows[r0 + 3:len(rows)]      # (the first samples still see the ramp)
    mhz = np.median([float(x[0]) for x in rs]); watts = np.median([float(x[1]) for x in rs])
    cap = sum(x[2].strip().lower().startswith("active") for x in rs)
    body = prof.get("body_fwd_tc")
    print("%-46s %4.0f MHz  %4.0f W  power-capped in %d of %d samples%s" % (
        label, mhz, watts, cap, len(rs), "  vertex kernel %.1f us" % (body[0] / body[1] * 1e3) if body else ""), flush=True)
    time.sleep(1.0)

probe("forward with verts (vertex kernel dominant)", lambda: smpl.forward_into(d["beta"], d["theta"], B, v, j, r))
if not os.environ.get("PROBE_VERTEX_ONLY"):
    probe("keypoint chain only (step without verts)", lambda: smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], want_verts=False, out=out))
proc.kill()
