"""GPU tests of the sharded step: the batch split over ranks, the visibility
count exchanged at the start of the step and the loss numerators next to the backward, through
(a) the mailbox kernels over peer memory and (b) NCCL.  The reference is single-device
(src/trainer.py:352); what has to hold is that a sharded step returns exactly the single-device
numbers: the loss of kp_reprojection_loss (src/ops.py:35-47) over the WHOLE batch and gradients that
divide by the GLOBAL num_present.

On one GPU the ranks are contexts of this process (smplb_comm_p2p_attach_local) that step ONE AFTER THE OTHER
(B200_PROFILING.md: kernels that wait for one another must not be separate launches on one GPU); the concurrent
exchange over CUDA IPC + NVLink is the 2-GPU torchrun test, skipped on a single-GPU box (logs in profiles/r02).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from hpe_b200 import runtime, synthetic
from hpe_b200.tf_smpl.batch_smpl import SMPL

pytestmark = pytest.mark.gpu
KEYS = ("joints", "kp_pred", "d_beta", "d_theta", "d_cam")


def _dev(ctx, inp, lo, hi):
    return {k: ctx.to_device(v[lo:hi]) for k, v in inp.items()}


def _np(out, keys=KEYS + ("loss_parts",)):
    return {k: out[k].numpy() for k in keys}


def _single_device_reference(model, inp, B, w_kp):
    s = SMPL(model, max_batch=B)
    full = s.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], w_kp=w_kp, want_verts=False)
    full = {k: np.array(v) for k, v in full.items() if v is not None}
    return s, full


def _run_ranks_in_turn(ranks, epoch, order, calls, local_parts):
    """One sharded step on every rank of a single GPU WITHOUT two kernels ever waiting for each other (nothing
    guarantees that two spinning kernels of one GPU run at the same time): the ranks step one after the other in
    `order`; what a rank needs from ranks that have not stepped yet is written into its mailbox with the test hook
    (their visibility count and numerators, computed beforehand without a communicator), what it needs from ranks
    that already stepped are those ranks' real pushes."""
    outs = [None] * len(ranks)
    done = set()
    for r in order:
        for peer in range(len(ranks)):
            if peer != r and peer not in done:
                cnt, num, mesh = local_parts[peer]
                ranks[r].ctx.p2p_inject(0, epoch, peer, cnt=cnt)
                ranks[r].ctx.p2p_inject(1, epoch, peer, v0=num, v1=mesh)
        outs[r] = calls[r]()
        ranks[r].ctx.sync()
        done.add(r)
    return outs


@pytest.mark.parametrize("nranks,B", [(2, 512), (3, 389)])
def test_mailbox_exchange_equals_single_device(full_model, nranks, B):
    """nranks contexts of one GPU act as ranks: every rank's loss is the whole batch's loss bit for bit
    on all ranks, the exchanged count is exact, and per-sample gradients equal the single-device ones
    computed with the global count (kp_count_override)."""
    from hpe_b200 import sharding
    w_kp = 60.0
    inp = synthetic.make_inputs(B, seed=77)
    ref_ctx, full = _single_device_reference(full_model, inp, B, w_kp)
    tot = int(full["loss_parts"][1])
    ranks = [SMPL(full_model, max_batch=B) for _ in range(nranks)]
    ctxs = [r.ctx for r in ranks]
    for i, r in enumerate(ranks):
        r.ctx.debug_set("comm_timeout_ms", 3000)
        r.ctx.p2p_attach_local(i, ctxs)
    rng = [sharding.shard_range(B, nranks, i) for i in range(nranks)]
    dev = [_dev(r.ctx, inp, *rng[i]) for i, r in enumerate(ranks)]
    # what every shard contributes, and the single-device per-sample results with the global count
    want, local_parts = [], []
    for lo, hi in rng:
        w = ref_ctx.step(inp["beta"][lo:hi], inp["theta"][lo:hi], inp["cam"][lo:hi], inp["kp_gt"][lo:hi], w_kp=w_kp,
                         want_verts=False, kp_count_override=tot)
        want.append({k: np.array(w[k]) for k in KEYS})
        local_parts.append((2 * int(np.count_nonzero(inp["kp_gt"][lo:hi, :, 2])), float(w["loss_parts"][0]), 0.0))
    calls = [lambda i=i: ranks[i].step(dev[i]["beta"], dev[i]["theta"], dev[i]["cam"], dev[i]["kp_gt"], w_kp=w_kp,
                                       want_verts=False) for i in range(nranks)]
    for epoch in (1, 2, 3, 4, 5, 6):          # more epochs than mailbox slots; the order the ranks step in rotates
        order = [(epoch + j) % nranks for j in range(nranks)]
        got = [_np(o) for o in _run_ranks_in_turn(ranks, epoch, order, calls, local_parts)]
        assert all(r.ctx.comm_status() == 0 for r in ranks)
        for g in got[1:]:
            assert np.array_equal(g["loss_parts"], got[0]["loss_parts"]), "ranks disagree on the global loss"
        lp = got[0]["loss_parts"]
        assert int(lp[1]) == tot
        assert abs(lp[0] - full["loss_parts"][0]) <= 2e-6 * full["loss_parts"][0]
        assert abs(lp[3] - full["loss_parts"][3]) <= 2e-6 * abs(full["loss_parts"][3])
        for i in range(nranks):
            for k in KEYS:
                assert np.array_equal(got[i][k], want[i][k]), "rank %d: %s differs from the single-device result" % (i, k)
    for r in ranks:
        r.ctx.comm_destroy()


def test_mailbox_exchange_mesh_step(full_model):
    """The mesh-loss step (keypoint + mesh numerators exchanged together)."""
    from hpe_b200 import ops, sharding
    B = 6
    inp = synthetic.make_inputs(B, seed=5)
    seg = synthetic.make_silhouettes(B, seed=6, a_range=(6, 10), b_range=(9, 14))
    pts3 = synthetic.silhouette_points(seg)
    s = SMPL(full_model, max_batch=B)
    full = s.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], silhouette=ops.silhouette_csr(pts3, B))
    full = {k: np.array(v) for k, v in full.items()}
    ranks = [SMPL(full_model, max_batch=B) for _ in range(2)]
    for i, r in enumerate(ranks):
        r.ctx.debug_set("comm_timeout_ms", 3000)
        r.ctx.p2p_attach_local(i, [x.ctx for x in ranks])
    calls, local_parts = [], []
    for i, r in enumerate(ranks):
        lo, hi = sharding.shard_range(B, 2, i)
        rows = pts3[(pts3[:, 0] >= lo) & (pts3[:, 0] < hi)].copy()
        rows[:, 0] -= lo
        pts, offs = ops.silhouette_csr(rows, hi - lo)
        d = _dev(r.ctx, inp, lo, hi)
        sil = (r.ctx.to_device(pts), r.ctx.to_device(offs, dtype=np.int32))
        calls.append(lambda r=r, d=d, sil=sil: r.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], silhouette=sil))
        w = s.step(inp["beta"][lo:hi], inp["theta"][lo:hi], inp["cam"][lo:hi], inp["kp_gt"][lo:hi], silhouette=(pts, offs))
        local_parts.append((int(w["loss_parts"][1]), float(w["loss_parts"][0]), float(w["loss_parts"][2])))
    for epoch, order in ((1, [0, 1]), (2, [1, 0])):
        got = [_np(o, ("loss_parts", "d_beta", "d_theta", "d_cam")) for o in _run_ranks_in_turn(ranks, epoch, order, calls, local_parts)]
        assert all(r.ctx.comm_status() == 0 for r in ranks)
        assert np.array_equal(got[0]["loss_parts"], got[1]["loss_parts"])
        for j in range(4):
            assert abs(got[0]["loss_parts"][j] - full["loss_parts"][j]) <= 1e-5 * abs(full["loss_parts"][j])
        for k in ("d_beta", "d_theta", "d_cam"):
            cat = np.concatenate([got[0][k], got[1][k]])
            assert np.max(np.abs(cat - full[k])) <= 1e-5 * np.max(np.abs(full[k])), k


def _three_context_stress(engines, B, steps, ref_engine=None):
    """Steps rotate over the contexts and over 4 input sets; every result must be bit-identical to the first
    one seen for the same (context, set) -- and to `ref_engine`'s (no communicator) when given."""
    NSET = 4
    host = [synthetic.make_inputs(B, seed=300 + i) for i in range(NSET)]
    dev = [[{k: e.ctx.to_device(v) for k, v in s.items()} for s in host] for e in engines]
    want = {}
    if ref_engine is not None:
        for s in range(NSET):
            o = ref_engine.step(host[s]["beta"], host[s]["theta"], host[s]["cam"], host[s]["kp_gt"], w_kp=60.0, want_verts=False)
            want[s] = {k: np.array(o[k]) for k in KEYS + ("loss_parts",)}
    first, bad, pending = {}, 0, []
    NE = len(engines)

    def check(item):
        nonlocal bad
        key, o, e = item
        engines[e].ctx.sync()
        got = {k: o[k].numpy() for k in KEYS + ("loss_parts",)}
        if want:
            for k in got:
                if not np.array_equal(got[k], want[key[1]][k]):
                    bad += 1
                    return
        blob = b"".join(v.tobytes() for v in got.values())
        if first.setdefault(key, blob) != blob:
            bad += 1

    for i in range(steps):
        e, s = i % NE, (i // NE) % NSET
        if len(pending) >= 4 * NE:
            check(pending.pop(0))
        d = dev[e][s]
        o = engines[e].step(d["beta"], d["theta"], d["cam"], d["kp_gt"], w_kp=60.0)
        pending.append(((e, s), o, e))
    while pending:
        check(pending.pop(0))
    return bad


@pytest.mark.parametrize("backend", ["mailbox", "nccl"])
def test_one_rank_communicator_three_contexts(full_model, backend):
    """The comm path of smplb_step with a 1-rank communicator: 3 contexts in flight, 300 steps, the
    low-priority GEMM stream on -- every step bit-equal to the path without a communicator.  (Round 1
    lost updates here: the all-reduce ran on the main stream, unordered against the reduction on stream3.)"""
    B = 1024
    engines = [SMPL(full_model, max_batch=B) for _ in range(3)]
    for e in engines:
        if backend == "mailbox":
            e.ctx.p2p_attach_local(0, [e.ctx])
        else:
            try:
                e.ctx.comm_init(1, 0, runtime.Context.comm_unique_id())
            except Exception as ex:    # libnccl missing on the box
                pytest.skip("NCCL unavailable: %s" % ex)
    ref = SMPL(full_model, max_batch=B)
    assert _three_context_stress(engines, B, 300, ref_engine=ref) == 0
    for e in engines:
        e.ctx.comm_destroy()


def test_exchange_timeout_poisons_loss_instead_of_hanging(full_model):
    """A peer that never arrives: the waiting kernels give up after the timeout, the loss is NaN and
    smplb_comm_status reports it -- the GPU is not left spinning."""
    B = 16
    a, b = SMPL(full_model, max_batch=B), SMPL(full_model, max_batch=B)
    for i, r in enumerate((a, b)):
        r.ctx.debug_set("comm_timeout_ms", 50)
        r.ctx.p2p_attach_local(i, [a.ctx, b.ctx])
    inp = synthetic.make_inputs(B, seed=9)
    out = a.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])     # rank 1 never steps
    assert np.isnan(out["loss_parts"][3]) and a.ctx.comm_status() == 1
    a.ctx.comm_destroy()
    b.ctx.comm_destroy()


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("backend", ["mailbox", "nccl"])
def test_two_gpu_torchrun_sharded_step(backend):
    """Two processes, two GPUs: tools/determinism.py (3 contexts in flight per rank, every step's global loss
    and gradients bit-equal to the first result for the same inputs AND to the single-process result with
    the global count) over CUDA IPC mailboxes / NCCL."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, SMPLB_COMM=backend)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "determinism.py"),
                        "240", "3", "2048"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "0 mismatches" in r.stdout
