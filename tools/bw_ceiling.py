"""Write-only (memset) and copy ceilings of the box, to put the kernels' GB/s in context."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpe_b200
from hpe_b200 import synthetic, runtime
from hpe_b200._lib import check, lib
from hpe_b200.tf_smpl.batch_smpl import SMPL
smpl = SMPL(synthetic.make_model(seed=0, num_verts=200, regressor_nnz=8), max_batch=1)
ctx = smpl.ctx
for mb in (340, 1024, 4096):
    n = mb << 20
    ctx.flush_l2(n); ctx.sync()
    ctx.timer_start(0)
    for _ in range(10):
        ctx.flush_l2(n)
    ctx.timer_stop(0)
    ms = ctx.timer_ms(0) / 10
    print("memset %5d MB: %.1f us  -> %.0f GB/s write-only" % (mb, ms * 1e3, n / ms / 1e6))
