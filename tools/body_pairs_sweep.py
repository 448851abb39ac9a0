"""How the fused blend+skinning kernel and the whole step react to giving the kernel fewer SM pairs
(the kernel is bound by L2 throughput, not by the SM count).  Usage: python tools/body_pairs_sweep.py"""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pairs in (74, 70, 66, 62, 56, 48):
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "300", "--debug", "body_pairs=%d" % pairs],
                         capture_output=True, text=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    print("pairs %2d: %.2f M meshes/s (%.1f us/step), body kernel alone %.1f us, e2e %.2f M" % (
        pairs, d["value"] / 1e6, d["ms_per_step"] * 1e3, d["kernels_ms_per_step"]["body_fwd_tc"] * 1e3, d["e2e"]["value"] / 1e6), flush=True)
