import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import hpe_b200
from hpe_b200 import ops, synthetic, runtime
B, V = 512, 6890
seg = synthetic.make_silhouettes(B, seed=2000)
pts3 = synthetic.silhouette_points(seg)
rng = np.random.default_rng(1)
base = (rng.normal(size=(B, V, 2)) * np.array([24.0, 43.0]) + 112.0).astype(np.float32)
ctx = ops._ctx_for(base)
pts, offs = ops.silhouette_csr(pts3, B)
dp, do = ctx.to_device(pts), ctx.to_device(offs, np.int32)
for label, sp in (("gaussian cloud like config 3", base), ("all inside (x0.3 about the centre)", ((base - 112.0) * 0.3 + 112.0).astype(np.float32)),
                  ("far: cloud shifted 60 px", base + np.array([60.0, 0.0], dtype=np.float32))):
    dsp = ctx.to_device(sp)
    for it in range(2): ops._mesh_call(ctx, dp, do, dsp, True, False)
    ctx.sync(); ctx.profile(True)
    for it in range(5): ops._mesh_call(ctx, dp, do, dsp, True, False)
    prof = ctx.profile_read(); ctx.profile(False)
    print(label, {k: round(ms / n, 3) for k, (ms, n) in prof.items() if "nn" in k or "build" in k})
