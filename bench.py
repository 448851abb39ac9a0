"""bench.py -- SMPL fwd + bwd + reprojection loss, meshes/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload of the headline (config.workload): BASELINE config 2 -- SMPL forward + keypoint projection
+ masked-L1 keypoint loss + backward w.r.t. beta/theta/cam at batch 4096 per GPU on synthetic
SMPL-topology constants (V=6890, dense 24-wide skinning weights, the real parent tree).  One
"step" = one such pass over one batch.  N>1 (torchrun, one process per GPU): every rank runs its own
4096-mesh shard (weak scaling); the only exchange -- the visibility count at the start of the step, the
loss numerators next to the backward -- happens inside the step's own kernels through mailboxes in peer
memory (SMPLB_COMM=nccl selects the NCCL all-reduce instead).

`value`     whole-job meshes/s with inputs resident in HBM: the MEDIAN over R blocks of exactly K steps
            each, every block bracketed by CUDA events (max over ranks per block); R is chosen so that
            the timed region lasts >= ~0.5 s whatever K is, and nvidia-smi samples the clocks inside it.
`e2e`       the same through the public host-buffer call: pinned host inputs -> H2D -> step -> D2H of
            loss + gradients, every step, inside the timed region (verts stay in HBM, as a trainer
            keeps them; `e2e_with_verts` also copies the 339 MB of verts back, PCIe-bound).
`roofline`  the dominant kernel's algorithmic bytes / its average launch time (events around every
            launch, that kernel alone on the GPU, after the warm-up: the conditions of the burst peak it is
            quoted against); `roofline_after_sustained` the same pass right after the >= 0.5 s region (the
            launch is power-capped: the clocks are lower then); `roofline_in_step` the same kernel timed
            inside the overlapped 3-context schedule of the timed region.
`configs`   the other BASELINE configs in the same line: c3 (B=1024, + mesh-reprojection loss + gradient
            penalty), c4 (strong scaling, global batch 32768 over the N ranks), c5 (inference sweep
            B = 1 ... 65536); c3 / c5 are single-GPU configs and run on rank 0 of the N=1 run only.
`cpu_baseline`  the reference formulation on the host cores (rank 0, N=1): the reference's own files
            under the torch-CPU `tensorflow` shim where the checkout exists, else the numpy port.
`--impl reference` times that CPU arm alone.
"""
import argparse
import json
import math
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
# SURVEY.md section 8d / BASELINE.md section 3: algorithmic bytes per mesh
BYTES_PER_MESH_STEP = 84832
BYTES_PER_MESH_INFER = 84400
CONST_BYTES = 19870760
WORKLOAD = ("SMPL fwd+bwd (beta/theta/cam) + kp reprojection loss, B=%d per GPU, V=6890, K=19, dense skinning weights "
            "(BASELINE config 2)")


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

    def mark(self):
        return len(self.rows)

    def window(self, first_row, last_row=None):
        """Clocks over the samples [first_row, last_row): the ones taken inside a timed region."""
        sm, mx, reasons = [], [], set()
        for r in self.rows[first_row:last_row]:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no nvidia-smi sample inside the region"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass


# ------------------------------------------------------------------------------- CPU reference arm
def _cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()


def make_cpu_arms(sample_batch):
    """[(kind, description, one_step)] -- forward + keypoint loss + backward of the reference formulation on
    `sample_batch` meshes: (a) "reference": the reference's own src/tf_smpl/*.py and src/ops.py executed under
    oracle/tf_shim (torch-CPU, fp32, gradients by autograd), only where the checkout exists (/root/reference
    in the build container; it cannot travel to the GPU box and is not pip-installable: no setup.py);
    (b) "port": oracle/smpl_numpy.py, numpy fp32 with hand-derived backward."""
    import hpe_b200  # noqa: F401
    from hpe_b200 import synthetic
    from oracle import smpl_numpy as onp
    model = synthetic.make_model(seed=0)
    inp = synthetic.make_inputs(sample_batch, seed=1000)
    arms = []
    try:
        import torch
        torch.set_num_threads(_cores())      # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host core
    except Exception:
        pass
    try:
        from oracle import run_reference
        for root in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
            if os.path.isfile(os.path.join(root, "src", "tf_smpl", "batch_smpl.py")):
                run_reference.REFERENCE_ROOT = root
                ref = run_reference.Reference(float64=False)
                smpl = ref.load_smpl(model)
                kp_gt = ref.tensor(inp["kp_gt"])

                def one_ref(ref=ref, smpl=smpl, kp_gt=kp_gt):
                    beta, theta, cam = (ref.tensor(inp[k], True) for k in ("beta", "theta", "cam"))
                    verts, joints, Rs = smpl(beta, theta, get_skin=True)
                    kp = ref.projection.batch_orth_proj_idrot(joints, cam)
                    loss = ref.ops.kp_reprojection_loss(kp_gt, kp)
                    ref.torch.autograd.grad(loss, [beta, theta, cam])

                arms.append(("reference", "the reference's own src/tf_smpl + src/ops.py under the torch-CPU tensorflow shim "
                                          "(fp32, autograd backward, %d torch threads)" % ref.torch.get_num_threads(), one_ref))
                break
    except Exception as ex:   # no checkout / no torch: the port stands alone
        sys.stderr.write("bench: reference checkout not usable (%s); timing the numpy port\n" % ex)
    o = onp.SMPL(model, dtype=np.float32)

    def one_port():
        verts, joints, Rs = o(inp["beta"], inp["theta"], get_skin=True)
        kp = onp.batch_orth_proj_idrot(joints, inp["cam"])
        onp.kp_reprojection_loss(inp["kp_gt"], kp)
        dj, dcam = onp.orth_proj_backward(joints, inp["cam"], onp.kp_loss_backward(inp["kp_gt"], kp))
        onp.smpl_backward(o, inp["beta"], inp["theta"], None, dj, None)

    arms.append(("port", "numpy fp32 port of the reference formulation (oracle/smpl_numpy.py; TensorFlow unavailable "
                         "offline, the reference checkout does not travel to the GPU box)", one_port))
    try:
        import torch
        from oracle import smpl_torch as ot
        ts = ot.SMPL(model, dtype=torch.float32)

        def one_torch():
            ot.step(ts, inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])

        arms.append(("port", "torch-CPU fp32 port of the reference formulation, line by line incl. its materialised weight tile "
                             "and 4x4 transforms (oracle/smpl_torch.py, autograd backward, %d torch threads; TensorFlow "
                             "unavailable offline, the reference checkout does not travel to the GPU box)"
                     % torch.get_num_threads(), one_torch))
    except Exception as ex:
        sys.stderr.write("bench: torch port not usable (%s)\n" % ex)
    return arms


def time_cpu_arm(one, steps, warmup, max_seconds=None):
    for _ in range(warmup):
        one()
    t0 = time.time()
    n = 0
    while n < steps:
        one()
        n += 1
        if max_seconds is not None and time.time() - t0 > max_seconds:
            break
    return (time.time() - t0) / n, n


def cpu_baseline(max_seconds, sample_batch=64):
    """The faster of the available CPU arms, each timed for ~max_seconds / #arms."""
    arms = make_cpu_arms(sample_batch)
    best, rates = None, {}
    for kind, desc, one in arms:
        dt, n = time_cpu_arm(one, 10 ** 9, 1, max_seconds / len(arms))
        label = kind + ":" + desc.split(" ")[0]
        rates[label] = sample_batch / dt
        if best is None or rates[label] > best["value"]:
            best = {"value": rates[label], "unit": UNIT, "cores": _cores(), "kind": kind,
                    "sample": "%d x %d meshes (chunks of the B=%d workload), fwd + kp loss + bwd: %s" % (n, sample_batch, BATCH, desc)}
    best["arms"] = rates
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 64
    arms = make_cpu_arms(sample)
    results = {}
    for kind, desc, one in arms:
        dt, n = time_cpu_arm(one, args.steps, max(args.warmup, 1), max_seconds=90.0 / len(arms))
        results[kind + ":" + desc.split(" ")[0]] = (sample / dt, dt, n, desc)
    label = max(results, key=lambda k: results[k][0])
    kind = label.split(":")[0]
    val, dt, n, desc = results[label]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
            "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % BATCH, "sample_batch": sample,
                       "note": "each step = one %d-mesh chunk of the workload on the host cores; the faster of the available "
                               "CPU arms is reported" % sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": _cores(), "kind": kind, "sample": desc,
                             "arms": {k: v[0] for k, v in results.items()}},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def repeat_until(fn, sync, min_ms=50.0, max_iter=100000):
    """Runs fn() in growing batches until one batch lasts >= min_ms; returns (ms per call, calls) of that batch."""
    n = 1
    while True:
        sync()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        sync()
        ms = (time.perf_counter() - t0) * 1e3
        if ms >= min_ms or n >= max_iter:
            return ms / n, n
        n = max(n * 2, int(n * min_ms / max(ms, 1e-3)) + 1)


def single_gpu_configs(smpl, model, peaks, sampler):
    """BASELINE configs 3 and 5 on one GPU (every measurement lasts >= 50 ms)."""
    from hpe_b200 import ops, synthetic
    from hpe_b200.tf_smpl.batch_smpl import SMPL
    ctx = smpl.ctx
    out = {}
    # ---- config 3: full HMR loss step at B = 1024: keypoint + mesh-reprojection loss + gradient penalty
    B3 = 1024
    inp = synthetic.make_inputs(B3, seed=3000)
    seg = synthetic.make_silhouettes(B3, seed=3001)
    d_seg = ctx.to_device(seg.reshape(B3, 224, 224))
    pts, offs = ops.silhouette_csr_device(d_seg, cap=int(seg.sum()) + 16)        # (for the brute-force comparison below)
    offs_h = offs.numpy()
    P = int(offs_h[-1])
    d = {k: ctx.to_device(v) for k, v in inp.items()}
    gp_in = [ctx.to_device(g) for g in synthetic.make_gp_inputs(3 * B3, seed=3002)]
    o3, ogp = {}, {}

    def c3_step():
        # the dense mask goes in as the trainer holds it: where(seg > 0) runs on the device inside the call, every step
        smpl.step(d["beta"], d["theta"], d["cam"], d["kp_gt"], seg=d_seg, w_kp=60.0, w_mesh=0.001, out=o3)
        ops.gradient_penalty_step(gp_in, out=ogp)

    for _ in range(3):
        c3_step()
    r0 = sampler.mark()
    ms3, n3 = repeat_until(c3_step, ctx.sync, 300.0)
    ck = sampler.window(r0, sampler.mark())
    ctx.profile(True)
    for _ in range(5):
        c3_step()
    prof = ctx.profile_read()
    ctx.profile(False)
    kms = {k: v[0] / 5 for k, v in prof.items()}
    mesh_ms = sum(v for k, v in kms.items() if k.startswith("mesh_"))
    gp_ms = sum(v for k, v in kms.items() if k.startswith("gp_"))
    pairs = 2.0 * P * V          # what the reference's scan evaluates: every (pixel, vertex) pair, both directions
    mhz = ck["sm_mhz"] or 1327.0
    fma_peak = 148 * 128 * mhz * 1e6
    # the reference's own algorithm (brute-force scan, 7 instructions per pair) on a 64-image subset
    sub = 64
    ctx.debug_set("mesh_grid", 0)
    sp = ctx.to_device(np.random.default_rng(1).uniform(40, 180, size=(sub, V, 2)).astype(np.float32))
    d_pts = ctx.to_device(pts.numpy()[:offs_h[sub]])
    d_offs = ctx.to_device(offs_h[:sub + 1], dtype=np.int32)
    ms_bf, _ = repeat_until(lambda: ops._mesh_call(ctx, d_pts, d_offs, sp, False, False), ctx.sync, 100.0)
    ctx.debug_set("mesh_grid", 1)
    pairs_bf = 2.0 * int(offs_h[sub]) * V
    gp_bytes = 3 * B3 * 428 * 4
    out["c3"] = {"workload": "BASELINE config 3: kp + mesh-reprojection loss + backward at B=1024 from the dense masks [B,224,224] "
                             "(where(seg > 0) on the device inside the step: P=%d silhouette pixels, %.0f per image) + critic "
                             "gradient penalty (forward + backward) over M=%d rows" % (P, P / B3, 3 * B3),
                 "ms_per_step": ms3, "value": B3 / (ms3 * 1e-3), "unit": UNIT, "iterations": n3, "clocks": ck,
                 "kernels_ms_per_step": kms,
                 "mesh_search": {"ms": mesh_ms, "reference_pairs_per_step": pairs,
                                 "reference_equivalent_pairs_per_s": pairs / (mesh_ms * 1e-3) if mesh_ms else None,
                                 "note": "pixel -> vertex on a binned grid, vertex -> pixel on the silhouette bitmap, both returning the "
                                         "brute-force scan's indices bit for bit; pairs counted as the reference's full scan "
                                         "(ops.py:60-71)"},
                 "mesh_brute_force": {"images": sub, "ms": ms_bf, "pairs_per_s": pairs_bf / (ms_bf * 1e-3),
                                      "fma_issue_peak_per_s": fma_peak,
                                      "frac_of_fma_issue_peak_at_7_instr_per_pair": 7 * pairs_bf / (ms_bf * 1e-3) / fma_peak},
                 "gradient_penalty": {"ms": gp_ms, "bytes": gp_bytes, "gbs": gp_bytes / (gp_ms * 1e-3) / 1e9 if gp_ms else None,
                                      "frac_of_hbm": gp_bytes / (gp_ms * 1e-3) / 1e9 / peaks["hbm"] if gp_ms else None}}
    for x in list(d.values()) + gp_in + [pts, offs, sp, d_pts, d_offs, d_seg]:
        x.free()
    # ---- config 5: inference sweep, SMPL forward (verts + joints + Rs), device-resident I/O
    sweep = []
    Bmax = 65536
    big = synthetic.make_inputs(Bmax, seed=5000)
    eng = SMPL(model, device=ctx.device, max_batch=Bmax)
    bctx = eng.ctx
    db, dt = bctx.to_device(big["beta"]), bctx.to_device(big["theta"])
    vb = bctx.empty((Bmax, V, 3))
    jb = bctx.empty((Bmax, K, 3))
    rb = bctx.empty((Bmax, 24, 3, 3))
    b = 1
    while b <= Bmax:
        fwd = lambda: eng.forward_into(db, dt, b, vb, jb, rb)       # noqa: E731
        for _ in range(3):
            fwd()
        ms, n = repeat_until(fwd, bctx.sync, 50.0)
        gbs = (BYTES_PER_MESH_INFER * b + CONST_BYTES) / (ms * 1e-3) / 1e9
        sweep.append({"B": b, "ms": ms, "meshes_per_s": b / (ms * 1e-3), "algorithmic_gbs": gbs, "frac_of_hbm": gbs / peaks["hbm"],
                      "iterations": n})
        b *= 2
    fj = lambda: eng.forward_into(db, dt, Bmax, None, jb, None)     # noqa: E731
    for _ in range(3):
        fj()
    msj, nj = repeat_until(fj, bctx.sync, 50.0)
    out["c5"] = {"workload": "BASELINE config 5: SMPL forward (verts, joints, Rs), device-resident I/O, B = 1 ... 65536; every point "
                             "timed for >= 50 ms",
                 "sweep": sweep, "latency_ms_B1": sweep[0]["ms"], "peak_meshes_per_s": max(s["meshes_per_s"] for s in sweep),
                 "joints_only_B65536": {"ms": msj, "meshes_per_s": Bmax / (msj * 1e-3)}}
    bctx.close()
    return out


# ------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--engines", type=int, default=3, help="contexts in flight per GPU")
    ap.add_argument("--region-steps", type=int, default=3000, help="steps of the whole timed region (R = ceil(this / K) blocks)")
    ap.add_argument("--no-extra", action="store_true", help="skip configs c3 / c4 / c5 and the CPU baseline (tuning runs)")
    ap.add_argument("--fused", type=int, default=-1, help="tuning: smplb_debug_set('fused', n) on every context")
    ap.add_argument("--debug", action="append", default=[], help="tuning: key=value for smplb_debug_set on every context")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    W = max(args.warmup, 3)
    KS = max(args.steps, 1)

    import hpe_b200  # noqa: F401
    from hpe_b200 import runtime, synthetic
    from hpe_b200.tf_smpl.batch_smpl import SMPL

    peaks = load_peaks()
    model = synthetic.make_model(seed=0)
    # Three contexts per GPU (own streams + workspace) take alternate steps, so the latency-bound
    # per-body kernels of one step fill the gaps of the other steps' streaming kernels.
    NE = max(1, args.engines)
    B_STRONG = 32768 // world
    engines = [SMPL(model, device=local, max_batch=max(B, B if args.no_extra else B_STRONG)) for _ in range(NE)]
    ctx = engines[0].ctx
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

    def sync_all():
        for e in engines:
            e.ctx.sync()

    def barrier():
        sync_all()
        if dist is not None:
            dist.barrier()
        sync_all()

    def max_over_ranks(values):
        if dist is None:
            return list(values)
        t = torch.tensor(list(values), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def timed_blocks(step, n_blocks, steps_per_block):
        """n_blocks blocks of exactly steps_per_block steps, each bracketed by CUDA events on context 0's stream
        (the other contexts fork from the start event and join before the stop event); no host
        synchronisation between blocks.  Returns the per-block ms (max over ranks)."""
        ms = []
        i = 0
        for r in range(n_blocks):
            slot = r % 16
            if r >= 16:
                ms.append(ctx.timer_ms(slot))            # block r - 16 finished long ago
            ctx.timer_start(slot)
            for e in engines[1:]:
                e.ctx.order_after(ctx)                   # the other contexts start after the start event
            for _ in range(steps_per_block):
                step(i)
                i += 1
            for e in engines[1:]:
                ctx.order_after(e.ctx)                   # the stop event waits for every context
            ctx.timer_stop(slot)
        for r in range(max(0, n_blocks - 16), n_blocks):
            ms.append(ctx.timer_ms(r % 16))
        return max_over_ranks(ms)

    # inputs: 4 rotating sets so no step re-reads what the previous one left in L2
    NSET = 4
    host_sets = [synthetic.make_inputs(B, seed=1000 + rank * 17 + i) for i in range(NSET)]
    dev_sets = [[{k: e.ctx.to_device(v) for k, v in s.items()} for s in host_sets] for e in engines]
    outs = [{} for _ in engines]

    def gpu_step(i, n_eng=NE):
        e = i % n_eng
        d = dev_sets[e][i % NSET]
        engines[e].step(d["beta"], d["theta"], d["cam"], d["kp_gt"], w_kp=60.0, out=outs[e])

    for i in range(NE * W):
        gpu_step(i)
    barrier()

    # ---- per-kernel launch durations: steps on ONE context with CUDA events around every launch (the library then
    # keeps all kernels on one stream so each duration is that kernel alone).  Taken twice: here, after the warm-up
    # (the conditions of MEASURED_PEAKS.json's burst figures, which `roofline` is quoted against), and again right after
    # the sustained region below, where the power cap holds the clocks lower (`roofline_after_sustained`).
    def kernel_pass():
        ctx.profile(True)
        for i in range(n_prof):
            gpu_step(i, 1)
        pr = ctx.profile_read()
        ctx.profile(False)
        return pr

    n_prof = min(KS, 200)
    prof = kernel_pass()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 8.0:      # nvidia-smi takes a while to start on an 8-GPU box
        time.sleep(0.05)
    # ---- headline: R blocks of exactly K steps; the region lasts >= ~0.5 s and the clocks are sampled inside it
    R = max(1, int(math.ceil(args.region_steps / float(KS))))
    launches0 = sum(e.ctx.launch_count() for e in engines)
    barrier()
    rows0 = sampler.mark()
    t_region = time.time()
    block_ms = timed_blocks(gpu_step, R, KS)
    barrier()
    t_region = time.time() - t_region
    rows1 = sampler.mark()
    clocks = sampler.window(rows0, rows1)
    launches = (sum(e.ctx.launch_count() for e in engines) - launches0) // R
    ms_block = float(np.median(block_ms))
    step_ms = ms_block / KS
    value = world * B * KS / (ms_block * 1e-3)
    loss_parts = outs[0]["loss_parts"].numpy()

    prof_hot = kernel_pass()
    # ... and inside the overlapped schedule of the timed region: events on the launching streams (trace mode)
    for e in engines:
        e.ctx.profile(2)
    n_trace = min(KS, 60)
    for i in range(NE * 2, NE * 2 + n_trace):
        gpu_step(i)
    trace = []
    for e in engines:
        trace += e.ctx.profile_trace()
        e.ctx.profile(False)
    barrier()

    # ---- e2e: host buffers in, loss + gradients out, copies inside the timed region.
    # The contexts (each with its own streams and workspace) alternate steps, so the PCIe copies
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

    E2E_STEPS = max(KS, 1500)
    for i in range(2 * NE):
        e2e_step(i)
    sync_all()
    barrier()
    t0 = time.perf_counter()
    for i in range(E2E_STEPS):
        e2e_step(i)
    sync_all()
    e2e_s = time.perf_counter() - t0
    e2e_ms = max_over_ranks([e2e_s * 1e3])[0]
    e2e_value = world * B * E2E_STEPS / (e2e_ms * 1e-3)
    h2d = sum(v.nbytes for v in host_sets[0].values())
    d2h = sum(eouts[0][k].nbytes for k in ("loss_parts", "d_beta", "d_theta", "d_cam"))
    # the same call returning verts to the host, as the reference's SMPL() does: PCIe-bound
    e2e_verts = None
    if not args.no_extra:
        vout = {"verts": runtime.pinned_empty((B, V, 3)), "loss_parts": eouts[0]["loss_parts"], "d_beta": eouts[0]["d_beta"],
                "d_theta": eouts[0]["d_theta"], "d_cam": eouts[0]["d_cam"]}
        p = pin[0]
        NV = 12
        for _ in range(2):
            engines[0].step(p["beta"], p["theta"], p["cam"], p["kp_gt"], w_kp=60.0, out=vout, skip=("Rs", "joints", "kp_pred"))
        barrier()
        t0 = time.perf_counter()
        for _ in range(NV):
            engines[0].step(p["beta"], p["theta"], p["cam"], p["kp_gt"], w_kp=60.0, out=vout, skip=("Rs", "joints", "kp_pred"))
        tv = max_over_ranks([(time.perf_counter() - t0) * 1e3])[0]
        e2e_verts = {"value": world * B * NV / (tv * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                     "d2h_bytes_per_step": int(d2h + vout["verts"].nbytes), "steps": NV,
                     "d2h_gbs_per_gpu": (d2h + vout["verts"].nbytes) * NV / (tv * 1e-3) / 1e9,
                     "note": "one context, synchronous host-mode calls that also copy verts [B,6890,3] back: PCIe-bound"}
        del vout

    extra = {}
    # ---- config 4: strong scaling, global batch 32768 split over the ranks (the same sharded step)
    if not args.no_extra:
        Bs = B_STRONG
        s_host = [synthetic.make_inputs(Bs, seed=4000 + rank * 17 + i) for i in range(2)]
        s_dev = [[{k: e.ctx.to_device(v) for k, v in s.items()} for s in s_host] for e in engines]
        s_outs = [{} for _ in engines]

        def strong_step(i):
            e = i % NE
            d = s_dev[e][i % 2]
            engines[e].step(d["beta"], d["theta"], d["cam"], d["kp_gt"], w_kp=60.0, out=s_outs[e])

        for i in range(2 * NE):
            strong_step(i)
        barrier()
        per_block = 48 * world                                # ~70 ms per block whatever N is
        ms = timed_blocks(strong_step, 5, per_block)
        barrier()
        mb = float(np.median(ms))
        extra["c4_strong_32768"] = {"workload": "BASELINE config 4: the config-2 step at global batch 32768 split over the ranks",
                                    "global_batch": 32768, "per_gpu_batch": Bs, "n_gpus": world,
                                    "ms_per_step": mb / per_block, "value": 32768 * per_block / (mb * 1e-3), "unit": UNIT,
                                    "scaling": "strong", "blocks": 5, "steps_per_block": per_block,
                                    "loss": float(s_outs[0]["loss_parts"].numpy()[3])}
        del s_dev, s_outs

    if world == 1 and not args.no_extra:
        extra.update(single_gpu_configs(engines[0], model, peaks, sampler))

    sampler.finish()
    line = None
    if rank == 0:
        # dominant kernel and its roofline
        dom = max(prof.items(), key=lambda kv: kv[1][0])
        name, (ms_sum, n) = dom
        avg_s = ms_sum / n * 1e-3
        work = kernel_work(name, B)
        roof = roof_in = roof_hot = None
        if work:
            bound, amount = work
            pk, unit, div = (peaks["hbm"], "GB/s", 1e9) if bound == "hbm" else (peaks["tf_sust"], "TFLOP/s", 1e12)
            ach = amount / avg_s / div
            roof = {"bound": bound, "achieved": ach, "peak": pk, "unit": unit, "frac": ach / pk, "traffic": None,
                    "kernel": name, "avg_launch_us": avg_s * 1e6, "peak_source": peaks["src"],
                    "timed": "events around every launch, %d steps on one context, one stream (the kernel alone), after the "
                             "warm-up and before the sustained region" % n_prof}
            if name in prof_hot:
                hs = prof_hot[name][0] / prof_hot[name][1] * 1e-3
                roof_hot = {"kernel": name, "avg_launch_us": hs * 1e6, "achieved": amount / hs / div, "frac": amount / hs / div / pk,
                            "unit": unit, "timed": "the same pass repeated right after the sustained region (clocks under the power cap)"}
            for tp in (os.path.join(ROOT, "profiles", "r02", "ncu_traffic.json"),
                       os.path.join(ROOT, "profiles", "r01", "ncu_traffic.json")):
                if os.path.isfile(tp) and B == BATCH:
                    roof["traffic"] = json.load(open(tp)).get(name)
                    break
            durs = [t1 - t0 for (nm, t0, t1) in trace if nm == name]
            if durs:
                a2 = amount / (float(np.mean(durs)) * 1e-3) / div
                roof_in = {"kernel": name, "avg_launch_us": float(np.mean(durs)) * 1e3, "achieved": a2, "frac": a2 / pk,
                           "unit": unit, "launches": len(durs),
                           "timed": "events on the launching stream inside the overlapped %d-context schedule" % NE}
        step_algo = (BYTES_PER_MESH_STEP * B + CONST_BYTES) / (step_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": KS, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD % B,
                       "global_batch": world * B, "parallelism": "batch-sharded x%d" % world, "contexts_in_flight_per_gpu": NE,
                       "exchange": comm_backend,
                       "l2": "per-step working set (verts, 340 MB per context) exceeds the 126 MB L2; inputs rotate "
                             "over %d buffer sets" % NSET,
                       "timing": "value = median of %d blocks of exactly %d steps, each block bracketed by CUDA events (max over "
                                 "ranks per block), steps rotating over the contexts in flight; the region of all blocks "
                                 "lasted %.2f s and the clocks were sampled inside it" % (R, KS, t_region),
                       "e2e": "host-buffer smpl.step rotating over the contexts in flight (copies of one step overlap kernels of "
                              "the other), %d steps: per step pinned H2D of beta/theta/cam/kp_gt and D2H of loss + d_beta/d_theta/"
                              "d_cam; verts are computed and stay in device memory (e2e_with_verts copies them back too)" % E2E_STEPS},
            "timed_blocks": R, "block_ms": {"median": ms_block, "min": float(min(block_ms)), "max": float(max(block_ms))},
            "timed_region_s": t_region,
            "roofline": roof, "roofline_after_sustained": roof_hot, "roofline_in_step": roof_in,
            "step_algorithmic_gbs": step_algo, "step_algorithmic_frac_of_hbm": step_algo / peaks["hbm"],
            "kernels_ms_per_step": {k: v[0] / n_prof for k, v in prof.items()},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": E2E_STEPS},
            "e2e_with_verts": e2e_verts,
            "gpu_launches": int(launches), "clocks": clocks,
            "loss": float(loss_parts[3]), "configs": extra,
        }
        if world == 1 and not args.no_extra:
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
    if dist is not None:
        # the global loss must be the same bits on every rank
        lt = torch.tensor([float(loss_parts[3])], device="cuda", dtype=torch.float64)
        lo, hi = lt.clone(), lt.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        st = max(e.ctx.comm_status() for e in engines)
        if rank == 0:
            line["global_loss_identical_on_all_ranks"] = bool(float(lo) == float(hi))
            line["exchange_status"] = int(st)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        for e in engines:
            e.ctx.comm_destroy()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
