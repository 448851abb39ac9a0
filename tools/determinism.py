"""Stress test for stream races: several contexts in flight take steps over rotating input sets, every
step's loss and gradients are copied out and must be bit-identical to the first result for the same
(context, input set).  Works under torchrun (the steps then all-reduce over ranks).
Usage: python tools/determinism.py [steps] [contexts] [B] [key=value,...  (smplb_debug_set)]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

import hpe_b200  # noqa: E402,F401
from hpe_b200 import runtime, synthetic  # noqa: E402
from hpe_b200.tf_smpl.batch_smpl import SMPL  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 600
NE = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
NSET = 4
model = synthetic.make_model(seed=0)
engines = [SMPL(model, device=local, max_batch=B) for _ in range(NE)]
for kv in (sys.argv[4].split(",") if len(sys.argv) > 4 else []):
    for e in engines:
        e.ctx.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
if world > 1:
    for e in engines:
        uid = [runtime.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        e.ctx.comm_init(world, rank, uid[0])
host_sets = [synthetic.make_inputs(B, seed=1000 + rank * 17 + i) for i in range(NSET)]
dev_sets = [[{k: e.ctx.to_device(v) for k, v in s.items()} for s in host_sets] for e in engines]
# one output set per step in flight so results can be read back later without serialising the steps
DEPTH = 4 * NE
outs = [{} for _ in range(DEPTH)]
ref = {}
bad = 0
pending = []


def check(item):
    global bad
    key, o, e = item
    engines[e].ctx.sync()
    got = (o["loss_parts"].numpy().tobytes(), o["d_theta"].numpy().tobytes(), o["d_beta"].numpy().tobytes(),
           o["d_cam"].numpy().tobytes())
    if key not in ref:
        ref[key] = got
    elif ref[key] != got:
        bad += 1
        if bad <= 5:
            names = ("loss", "d_theta", "d_beta", "d_cam")
            print("rank %d: MISMATCH context %d set %d in %s; loss_parts %s vs first %s" % (
                rank, key[0], key[1], [n for n, a, b in zip(names, ref[key], got) if a != b],
                np.frombuffer(got[0], np.float32), np.frombuffer(ref[key][0], np.float32)), flush=True)


for i in range(STEPS):
    e = i % NE
    s = (i // NE) % NSET
    o = outs[i % DEPTH]
    if len(pending) >= DEPTH:
        check(pending.pop(0))
    d = dev_sets[e][s]
    engines[e].step(d["beta"], d["theta"], d["cam"], d["kp_gt"], w_kp=60.0, out=o)
    pending.append(((e, s), o, e))
while pending:
    check(pending.pop(0))
print("rank %d: %d steps, %d contexts, %d mismatches" % (rank, STEPS, NE, bad), flush=True)
if dist is not None:
    dist.barrier()
sys.exit(1 if bad else 0)
