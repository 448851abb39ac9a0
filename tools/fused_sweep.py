"""Time the tile-shape variants of the fused blend+skinning kernel (k_body_tc.cu) and the
two-kernel path at one batch size.  Usage: python tools/fused_sweep.py [B]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200  # noqa: E402,F401
from hpe_b200 import synthetic  # noqa: E402
from hpe_b200.tf_smpl.batch_smpl import SMPL  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
smpl = SMPL(synthetic.make_model(seed=0), max_batch=B)
ctx = smpl.ctx
inp = synthetic.make_inputs(B, seed=1000)
d = {k: ctx.to_device(v) for k, v in inp.items()}
out = {}
names = {0: "two kernels", 1: "CTA pairs, resident Dt16 tile, W16 in TMEM, rings 3 / 4 (default)", 10: "the same with rings 2 / 3", 7: "CTA pairs, streaming Dt16", 9: "resident Dt16 tile, sixteen epilogue warps", 2: "NS=96 ST=8 x2 PRE=0", 3: "NS=128 ST=4 x2 PRE=8",
         4: "NS=96 ST=8 x2 PRE=4 (best single-CTA)", 5: "W16 in TMEM, vertex-tile major", 6: "CTA pairs, Dt16 multicast"}
for variant in ([int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (0, 1, 9, 7, 4)):
    ctx.debug_set("fused", variant)
    for it in range(3):
        smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], out=out)
    ctx.sync()
    ctx.profile(True)
    N = 10
    for it in range(N):
        smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], out=out)
    prof = ctx.profile_read()
    ctx.profile(False)
    heavy = sum(ms for k, (ms, n) in prof.items() if k in ("body_fwd_tc", "blend_fwd_tc", "skin_fwd_tc")) / N
    print("variant %d (%-40s): %.1f us" % (variant, names[variant], heavy * 1e3))
