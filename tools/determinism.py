"""Stress test for stream races: several contexts in flight take steps over rotating input sets, every
step's loss and gradients are copied out and must be bit-identical to the first result for the same
(context, input set).  Under torchrun the batch is sharded over the ranks (SMPLB_COMM=mailbox: CUDA IPC
mailboxes, the default; SMPLB_COMM=nccl); then the gradients must ALSO equal the single-process result
computed with the global visibility count, and the global loss must be the same bits on every rank.
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
BACKEND = os.environ.get("SMPLB_COMM", "mailbox")
host_sets = [synthetic.make_inputs(B, seed=1000 + rank * 17 + i) for i in range(NSET)]
want_grads = {}
if world > 1:
    # single-process reference with the global count, before any communicator is attached
    for s, h in enumerate(host_sets):
        cnt = torch.tensor([2 * int(np.count_nonzero(h["kp_gt"][:, :, 2]))], device="cuda", dtype=torch.int64)
        dist.all_reduce(cnt)
        o = engines[0].step(h["beta"], h["theta"], h["cam"], h["kp_gt"], w_kp=60.0, want_verts=False,
                            kp_count_override=int(cnt.item()))
        want_grads[s] = (np.array(o["d_theta"]).tobytes(), np.array(o["d_beta"]).tobytes(), np.array(o["d_cam"]).tobytes())
    for e in engines:
        if BACKEND == "nccl":
            uid = [runtime.Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            e.ctx.comm_init(world, rank, uid[0])
        else:
            hs = [None] * world
            dist.all_gather_object(hs, e.ctx.p2p_export())
            e.ctx.p2p_attach(world, rank, hs)
    dist.barrier()
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
    if want_grads and got[1:] != want_grads[key[1]]:
        bad += 1
        if bad <= 5:
            print("rank %d: context %d set %d: gradients differ from the single-process result with the global count"
                  % (rank, key[0], key[1]), flush=True)
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
if dist is not None:
    # the global loss of every (context, set) is the same bits on every rank
    mine = np.frombuffer(b"".join(ref[k][0] for k in sorted(ref)), np.float32).copy()
    t = torch.from_numpy(mine).cuda()
    all_t = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(all_t, t)
    for r, o in enumerate(all_t):
        if not torch.equal(o.view(torch.int32), t.view(torch.int32)):
            bad += 1
            print("rank %d: global loss differs from rank %d's" % (rank, r), flush=True)
    st = [e.ctx.comm_status() for e in engines]
    if any(st):
        bad += 1
        print("rank %d: exchange timed out %s" % (rank, st), flush=True)
print("rank %d: %d steps, %d contexts, backend %s, %d mismatches" % (rank, STEPS, NE, BACKEND if world > 1 else "none", bad),
      flush=True)
if dist is not None:
    dist.barrier()
    for e in engines:
        e.ctx.comm_destroy()
    dist.barrier()
sys.exit(1 if bad else 0)
