"""World-size-2 gloo test (CPU) of the N>1 host logic: shard the batch, reduce
{kp numerator, kp count, mesh sum, 428 GP column sums} with one all-reduce, and
recover exactly the single-process losses.  The per-rank compute here is the
oracle (this is a CPU test of the sharding recipe, not of the kernels)."""
import os
import socket

import numpy as np
import pytest

from hpe_b200 import sharding, synthetic
from oracle import smpl_numpy as onp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 9
    inp = synthetic.make_inputs(B, seed=321, dtype=np.float64)
    rng = np.random.default_rng(5)
    kp_pred = rng.uniform(-1, 1, size=(B, 19, 2))
    gp = synthetic.make_gp_inputs(3 * B, seed=8, dtype=np.float64)
    lo, hi = sharding.shard_range(B, world, rank)
    num, cnt = onp.kp_loss_parts(inp["kp_gt"][lo:hi], kp_pred[lo:hi])
    glo, ghi = sharding.shard_range(3 * B, world, rank)
    cols = np.concatenate([g[glo:ghi].reshape(ghi - glo, -1).sum(0) for g in gp])
    vec = torch.from_numpy(sharding.pack_partials(num, cnt, 0.0, cols))
    dist.all_reduce(vec)
    res = sharding.finish_losses(vec.numpy(), m_total=3 * B)
    local_loss = num / cnt if cnt else 0.0
    q.put((rank, res, local_loss, (lo, hi)))
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    for B in (1, 7, 64, 4096, 32768):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(B, w, i) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_gloo_allreduce_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B = 9
    inp = synthetic.make_inputs(B, seed=321, dtype=np.float64)
    kp_pred = np.random.default_rng(5).uniform(-1, 1, size=(B, 19, 2))
    want = onp.kp_reprojection_loss(inp["kp_gt"], kp_pred)
    want_gp = onp.compute_gradient_penalty(synthetic.make_gp_inputs(3 * B, seed=8, dtype=np.float64))
    for rank, res, local_loss, rng in got:
        assert res["kp_loss"] == pytest.approx(want, rel=1e-6)
        assert res["kp_count"] == 2 * int(np.count_nonzero(inp["kp_gt"][:, :, 2]))
        assert res["gradient_penalty"] == pytest.approx(want_gp, rel=1e-5)
    # the naive alternative (mean of per-shard losses) is NOT the reference's loss
    naive = np.mean([g[2] for g in got])
    assert abs(naive - want) > 1e-4 * want
