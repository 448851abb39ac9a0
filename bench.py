"""bench.py -- SMPL fwd + bwd + reprojection loss, meshes/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE config 2 -- SMPL forward + keypoint projection +
masked-L1 keypoint loss + backward w.r.t. beta/theta/cam at batch 4096 per GPU on synthetic
SMPL-topology constants (V=6890, dense 24-wide skinning weights, the real parent tree).  One
"step" = one such pass over one batch.  N>1 (torchrun, one process per GPU): every rank
runs its own 4096-mesh shard (weak scaling, no data-path collective); the only exchange is
the NCCL all-reduce of the loss numerator / visibility count.

`value`     whole-job meshes/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`       the same through the public host-buffer call: pinned host inputs -> H2D -> step
            -> D2H of loss + gradients, every step, inside the timed region.
`roofline`  the dominant kernel's algorithmic bytes (or flops) / its average launch time.
`cpu_baseline`  the numpy port of the reference formulation (oracle/) on the host cores.
`--impl reference` times that CPU port alone (TensorFlow, which the reference needs, is not
installable offline -- DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 4096
V = 6890
K = 19
METRIC = "smpl_fwd_bwd_reproj_loss_meshes_per_sec"
UNIT = "meshes/s"
# SURVEY.md §8d / BASELINE.md §3: algorithmic bytes per mesh of the whole step
BYTES_PER_MESH_STEP = 84832
CONST_BYTES = 19870760


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sust": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sust": 1400.0, "src": "fallback"}


# Per-kernel algorithmic work for one launch at batch B (DESIGN.md "kernels" table).
def kernel_work(name, B):
    vb = 12.0 * V   # bytes of one [V,3] fp32 row
    table = {
        # name: (bound, bytes or flops per launch)
        "pose_fwd": ("hbm", B * (340 + 864 + 288 + 1152 + 288 + 512 + 3072 + 1408)),
        "fold_gemm_u": ("tensor", 2.0 * B * 1368 * 218 * 3),
        "fold_gemm_dx": ("tensor", 2.0 * B * 1368 * 218 * 3),
        "fold_joints_proj_kploss": ("hbm", B * (1368 * 4 + 1152 + K * 44)),
        "fold_bwd_du_dA": ("hbm", B * (1368 * 4 + 1152 + K * 12 + 1152 + 3 * 1408 * 2)),
        "blend_fwd_sgemm": ("tensor", 2.0 * B * 218 * 3 * V),
        "blend_fwd_tc": ("hbm", B * (vb + 512) + 2 * 256 * 3 * V),
        # fused blend + skinning: verts out, x16 (512 B) + A16 (1536 B) rows in, Dt16 + W16 once
        "body_fwd_tc": ("hbm", B * (vb + 512 + 1536) + 2 * 256 * 3 * V + 2 * 64 * V),
        "skin_fwd": ("hbm", B * (2 * vb + 1152) + 24 * 4 * V),
        "skin_fwd_tc": ("hbm", B * (2 * vb + 3072) + 2 * 128 * V),
        "skin_bwd_active": ("hbm", B * (1152 + K * 12 + 4 * 1152 + 2 * 608 * 12)),
        "blend_bwd_sgemm_active": ("tensor", 2.0 * B * 217 * 3 * 608),
        "joints_proj_kploss": ("hbm", B * (K * 32 * 32 + K * 12 + K * 20 + 12)),
        "skin_bwd": ("hbm", B * (2 * vb + 1152 + K * 12 + 4 * 1152)),
        "blend_bwd_sgemm": ("tensor", 2.0 * B * 217 * 3 * V),
        "pose_bwd": ("hbm", B * (4 * 1152 + 16 * 4 * 224 + 864 + 288 + 1152 + 288 + 328)),
        "proj_bwd": ("hbm", B * (K * 32)),
        "reduce_kp": ("hbm", B * 8),
        "finalize_loss": ("hbm", 32),
    }
    return table.get(name)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self, first_row=0):
        """Clocks over the samples taken from row `first_row` on (the ones under load)."""
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows[first_row:]:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        busy = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_rate(max_seconds, sample_batch=64):
    """Forward + keypoint loss + backward with the numpy port of the reference's
    formulation (oracle/smpl_numpy.py, fp32), all host threads BLAS can use."""
    import hpe_b200  # noqa: F401
    from hpe_b200 import synthetic
    from oracle import smpl_numpy as onp
    model = synthetic.make_model(seed=0)
    o = onp.SMPL(model, dtype=np.float32)
    inp = synthetic.make_inputs(sample_batch, seed=1000)

    def one():
        verts, joints, Rs = o(inp["beta"], inp["theta"], get_skin=True)
        kp = onp.batch_orth_proj_idrot(joints, inp["cam"])
        loss = onp.kp_reprojection_loss(inp["kp_gt"], kp)
        dj, dcam = onp.orth_proj_backward(joints, inp["cam"], onp.kp_loss_backward(inp["kp_gt"], kp))
        db, dth = onp.smpl_backward(o, inp["beta"], inp["theta"], None, dj, None)
        return loss

    one()
    t0 = time.time()
    n = 0
    while True:
        one()
        n += 1
        if time.time() - t0 > max_seconds or n >= 200:
            break
    dt = time.time() - t0
    return sample_batch * n / dt, n, sample_batch


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    sample = 64
    import hpe_b200  # noqa: F401
    from hpe_b200 import synthetic
    from oracle import smpl_numpy as onp
    model = synthetic.make_model(seed=0)
    o = onp.SMPL(model, dtype=np.float32)
    inp = synthetic.make_inputs(sample, seed=1000)

    def one():
        verts, joints, Rs = o(inp["beta"], inp["theta"], get_skin=True)
        kp = onp.batch_orth_proj_idrot(joints, inp["cam"])
        onp.kp_reprojection_loss(inp["kp_gt"], kp)
        dj, dcam = onp.orth_proj_backward(joints, inp["cam"], onp.kp_loss_backward(inp["kp_gt"], kp))
        onp.smpl_backward(o, inp["beta"], inp["theta"], None, dj, None)

    for _ in range(args.warmup):
        one()
    t0 = time.time()
    for _ in range(args.steps):
        one()
    dt = time.time() - t0
    val = sample * args.steps / dt
    desc = ("numpy fp32 port of the reference formulation (oracle/smpl_numpy.py): fwd + kp loss + bwd on %d meshes per "
            "step, chunk of the B=%d workload; TensorFlow unavailable offline" % (sample, BATCH))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SMPL fwd+bwd (beta/theta/cam) + kp reprojection loss, B=%d per GPU, V=6890" % BATCH,
                       "sample_batch": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--engines", type=int, default=3, help="contexts in flight per GPU")
    ap.add_argument("--fused", type=int, default=-1, help="tuning: smplb_debug_set('fused', n) on every context")
    ap.add_argument("--debug", action="append", default=[], help="tuning: key=value for smplb_debug_set on every context")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    W = max(args.warmup, 3)

    import hpe_b200  # noqa: F401
    from hpe_b200 import runtime, synthetic
    from hpe_b200.tf_smpl.batch_smpl import SMPL

    peaks = load_peaks()
    model = synthetic.make_model(seed=0)
    # Two contexts per GPU (own streams + workspace) take alternate steps, so the latency-bound
    # per-body kernels of one step fill the gaps of the other step's streaming kernels.
    NE = max(1, args.engines)
    engines = [SMPL(model, device=local, max_batch=B) for _ in range(NE)]
    smpl = engines[0]
    ctx = smpl.ctx
    if args.fused >= 0:
        for e in engines:
            e.ctx.debug_set("fused", args.fused)
    for kv in args.debug:
        k, v = kv.split("=")
        for e in engines:
            e.ctx.debug_set(k, int(v))
    comm_backend = os.environ.get("SMPLB_COMM", "mailbox") if world > 1 else "none"
    if world > 1:
        # the batch is sharded over the ranks; smplb_step exchanges the visibility count and the loss
        # numerators inside its own kernels through mailboxes in peer memory (CUDA IPC handles passed
        # around with torch.distributed), or with NCCL when SMPLB_COMM=nccl
        for e in engines:
            if comm_backend == "nccl":
                uid = [runtime.Context.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                e.ctx.comm_init(world, rank, uid[0])
            else:
                hs = [None] * world
                dist.all_gather_object(hs, e.ctx.p2p_export())
                e.ctx.p2p_attach(world, rank, hs)
        dist.barrier()

    # inputs: 4 rotating sets so no step re-reads what the previous one left in L2
    NSET = 4
    host_sets = [synthetic.make_inputs(B, seed=1000 + rank * 17 + i) for i in range(NSET)]
    dev_sets = [[{k: e.ctx.to_device(v) for k, v in s.items()} for s in host_sets] for e in engines]
    outs = [{} for _ in engines]

    def sync_all():
        for e in engines:
            e.ctx.sync()

    def barrier():
        sync_all()
        if dist is not None:
            dist.barrier()
        sync_all()

    def gpu_step(i, n_eng=NE):
        e = i % n_eng
        d = dev_sets[e][i % NSET]
        # with world > 1 the context holds an NCCL communicator and smplb_step all-reduces
        # {kp numerator, kp count, mesh sum} inside the call (the path's one exchange)
        engines[e].step(d["beta"], d["theta"], d["cam"], d["kp_gt"], w_kp=60.0, out=outs[e])

    for i in range(NE * W):
        gpu_step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 8.0:      # nvidia-smi takes a while to start on an 8-GPU box
        time.sleep(0.05)
    rows0 = len(sampler.rows)
    launches0 = sum(e.ctx.launch_count() for e in engines)
    barrier()
    ctx.timer_start(0)
    for e in engines[1:]:
        e.ctx.order_after(ctx)                      # the other engines start after the start event
    for i in range(args.steps):
        gpu_step(i)
    for e in engines[1:]:
        ctx.order_after(e.ctx)                      # the stop event waits for every engine
    ctx.timer_stop(0)
    ms_total = ctx.timer_ms(0)
    barrier()
    launches = sum(e.ctx.launch_count() for e in engines) - launches0
    # per-kernel launch durations: K steps on ONE context with CUDA events around every launch
    # (the library then keeps all kernels on one stream so each duration is clean)
    ctx.profile(True)
    for i in range(args.steps):
        gpu_step(i, 1)
    prof = ctx.profile_read()
    ctx.profile(False)
    # the timed region is ~0.1 s: keep the same load running (untimed, the same number of steps on every
    # rank: the steps all-reduce) for another ~0.4 s so that nvidia-smi samples the clocks under it
    for i in range(2000):
        gpu_step(i)
    sync_all()
    barrier()
    clocks = sampler.finish(rows0)
    loss_parts = outs[0]["loss_parts"].numpy()

    # ---- e2e: host buffers in, loss + gradients out, copies inside the timed region.
    # Two contexts (each with its own streams and workspace) alternate steps, so the PCIe copies
    # of one step overlap the kernels of the other -- the double buffering any input pipeline
    # does.  Every step still copies ITS inputs H2D from pinned memory and ITS loss + gradients
    # D2H; verts (339 MB) and Rs are computed and stay in device memory.
    pin = [{k: runtime.pinned_empty(v.shape) for k, v in s.items()} for s in host_sets]
    for p, s in zip(pin, host_sets):
        for k in s:
            p[k][...] = s[k]
    eouts = [{"verts": None, "loss_parts": runtime.pinned_empty((4,)), "d_beta": runtime.pinned_empty((B, 10)),
              "d_theta": runtime.pinned_empty((B, 72)), "d_cam": runtime.pinned_empty((B, 3))} for _ in engines]
    losses = []

    def e2e_step(i):
        e = i % NE
        engines[e].ctx.sync()                       # results of step i-NE (same engine) are on the host now
        if i >= NE:
            losses.append(float(eouts[e]["loss_parts"][3]))
        p = pin[i % NSET]
        engines[e].step(p["beta"], p["theta"], p["cam"], p["kp_gt"], w_kp=60.0, want_verts="device", out=eouts[e],
                        skip=("Rs", "joints", "kp_pred"), nowait=True)

    for i in range(2 * NE):
        e2e_step(i)
    for e in engines:
        e.ctx.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    for e in engines:
        e.ctx.sync()
    e2e_s = time.perf_counter() - t0
    h2d = sum(v.nbytes for v in host_sets[0].values())
    d2h = sum(eouts[0][k].nbytes for k in ("loss_parts", "d_beta", "d_theta", "d_cam"))

    # ---- max over ranks
    if dist is not None:
        import torch
        t = torch.tensor([ms_total, e2e_s * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms = float(t[0]), float(t[1])
    else:
        e2e_ms = e2e_s * 1e3
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        # dominant kernel and its roofline
        dom = max(prof.items(), key=lambda kv: kv[1][0])
        name, (ms_sum, n) = dom
        avg_s = ms_sum / n * 1e-3
        work = kernel_work(name, B)
        roof = None
        if work:
            bound, amount = work
            if bound == "hbm":
                ach = amount / avg_s / 1e9
                roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                        "traffic": None}
            else:
                ach = amount / avg_s / 1e12
                roof = {"bound": "tensor", "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tf_sust"], "traffic": None}
            tpath = os.path.join(ROOT, "profiles", "r01", "ncu_traffic.json")
            if os.path.isfile(tpath) and B == BATCH:
                roof["traffic"] = json.load(open(tpath)).get(name)
            roof["kernel"] = name
            roof["avg_launch_us"] = avg_s * 1e6
            roof["peak_source"] = peaks["src"]
        step_ms = ms_total / args.steps
        e2e_algo = (BYTES_PER_MESH_STEP * B + CONST_BYTES) / (step_ms * 1e-3) / 1e9
        cpu_val, cpu_n, cpu_b = cpu_reference_rate(args.cpu_seconds)
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "SMPL fwd+bwd (beta/theta/cam) + kp reprojection loss, B=%d per GPU, V=6890, K=19, "
                                   "dense skinning weights (BASELINE config 2)" % B,
                       "global_batch": world * B, "parallelism": "batch-sharded x%d" % world, "contexts_in_flight_per_gpu": NE,
                       "l2": "per-step working set (verts, 340 MB per context) exceeds the 126 MB L2; inputs rotate "
                             "over %d buffer sets" % NSET,
                       "timing": "value: CUDA events around the K steps, which rotate over the contexts in flight on the GPU "
                                 "(each overlaps its 6890-vertex kernels with its keypoint path on a second stream); "
                                 "roofline / kernels_ms_per_step: K more steps on one context with events around every "
                                 "launch, single stream",
                       "e2e": "host-buffer smpl.step rotating over the contexts in flight (copies of one step overlap kernels of the "
                              "other): per step pinned H2D of beta/theta/cam/kp_gt and D2H of loss + d_beta/d_theta/d_cam; "
                              "verts are computed and stay in device memory"},
            "roofline": roof,
            "step_algorithmic_gbs": e2e_algo, "step_algorithmic_frac_of_hbm": e2e_algo / peaks["hbm"],
            "kernels_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d x %d meshes, fwd + kp loss + bwd, numpy fp32 port of the reference formulation "
                                       "(TensorFlow unavailable offline)" % (cpu_n, cpu_b)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clocks,
            "loss": float(loss_parts[3]),
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
