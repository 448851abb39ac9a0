"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes facade),
against (a) the committed golden vectors produced by the reference's own code
and (b) the fp64 numpy oracle on fresh seeded inputs.

Tolerance: fp32 kernels must satisfy
max|x - ref| <= 1e-4 * max|ref| per tensor against the fp64 oracle; masking
and counts are bit-exact.  Nothing here reads /root/reference.
"""
import numpy as np
import pytest

from conftest import rel_err
from hpe_b200 import ops, runtime, synthetic
from hpe_b200.tf_smpl import batch_lbs, projection
from hpe_b200.tf_smpl.batch_smpl import SMPL
from oracle import smpl_numpy as onp

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def smpl_small(small_model):
    return SMPL(small_model, max_batch=8)


@pytest.fixture(scope="module")
def smpl_full(full_model):
    return SMPL(full_model, max_batch=64)


def f64(d, *keys):
    return [d[k].astype(np.float64) for k in keys]


# ------------------------------------------------------------------ golden vectors
def test_forward_golden_small(smpl_small, golden_small):
    g = golden_small
    verts, joints, Rs = smpl_small(g["beta"], g["theta"], get_skin=True)
    assert verts.dtype == np.float32 and verts.shape == g["verts"].shape
    assert rel_err(verts, g["verts"]) < TOL
    assert rel_err(joints, g["joints"]) < TOL
    assert rel_err(Rs, g["Rs"]) < TOL
    assert rel_err(smpl_small.J_transformed, g["J_transformed"]) < TOL
    # get_skin=False returns joints only (batch_smpl.py:157-160)
    j2 = smpl_small(g["beta"], g["theta"])
    assert np.array_equal(j2, joints)
    # theta == 0 (sample 0): Rs is exactly the identity
    assert np.array_equal(Rs[0], np.broadcast_to(np.eye(3, dtype=np.float32), (24, 3, 3)))


def test_forward_golden_full(smpl_full, golden_full):
    g = golden_full
    st = int(g["vert_stride"])
    verts, joints, Rs = smpl_full(g["beta"], g["theta"], get_skin=True)
    assert rel_err(verts[:, ::st], g["verts_sub"]) < TOL
    assert rel_err(joints, g["joints"]) < TOL
    assert rel_err(Rs, g["Rs"]) < TOL
    assert rel_err(smpl_full.J_transformed, g["J_transformed"]) < TOL


def test_forward_lsp(small_model, golden_small_lsp):
    g = golden_small_lsp
    s = SMPL(small_model, joint_type="lsp", max_batch=4)
    joints = s(g["beta"], g["theta"])
    assert joints.shape == (3, 14, 3)
    assert rel_err(joints, g["joints"]) < TOL
    out = s.step(g["beta"], g["theta"], g["cam"], g["kp_gt"], w_kp=1.0)
    assert abs(out["loss_parts"][3] - float(g["kp_loss"])) < TOL * abs(float(g["kp_loss"]))
    assert rel_err(out["d_beta"], g["kp_d_beta"]) < TOL
    assert rel_err(out["d_theta"], g["kp_d_theta"]) < TOL
    assert rel_err(out["d_cam"], g["kp_d_cam"]) < TOL


@pytest.mark.parametrize("which", ["small", "full"])
def test_backward_upstream_golden(which, smpl_small, smpl_full, golden_small, golden_full):
    s, g = (smpl_small, golden_small) if which == "small" else (smpl_full, golden_full)
    s(g["beta"], g["theta"], get_skin=True)
    d_beta, d_theta = s.backward(g["up_verts"], g["up_joints"], g["up_Rs"])
    assert rel_err(d_beta, g["up_d_beta"]) < TOL
    assert rel_err(d_theta, g["up_d_theta"]) < TOL


@pytest.mark.parametrize("which", ["small", "full"])
def test_step_kp_golden(which, smpl_small, smpl_full, golden_small, golden_full):
    s, g = (smpl_small, golden_small) if which == "small" else (smpl_full, golden_full)
    out = s.step(g["beta"], g["theta"], g["cam"], g["kp_gt"], w_kp=1.0)
    assert rel_err(out["kp_pred"], g["kp_pred"]) < TOL
    lp = out["loss_parts"]
    # visibility count is bit-exact (integer) and the loss is numerator / count
    vis = g["kp_gt"][:, :, 2]
    assert int(lp[1]) == 2 * int(np.count_nonzero(vis))
    assert abs(lp[3] - float(g["kp_loss"])) < TOL * abs(float(g["kp_loss"]))
    assert rel_err(out["d_beta"], g["kp_d_beta"]) < TOL
    assert rel_err(out["d_theta"], g["kp_d_theta"]) < TOL
    assert rel_err(out["d_cam"], g["kp_d_cam"]) < TOL
    # sample 1 is all-invisible: its gradients are EXACTLY zero
    assert not out["d_beta"][1].any() and not out["d_theta"][1].any() and not out["d_cam"][1].any()


@pytest.mark.parametrize("which", ["small", "full"])
def test_step_kp_plus_mesh_golden(which, smpl_small, smpl_full, golden_small, golden_full):
    """60 * kp + 0.001 * mesh, the trainer's weighting (config.py:67-68)."""
    s, g = (smpl_small, golden_small) if which == "small" else (smpl_full, golden_full)
    B = g["beta"].shape[0]
    sil = ops.silhouette_csr(g["seg_points"], B)
    out = s.step(g["beta"], g["theta"], g["cam"], g["kp_gt"], silhouette=sil, w_kp=60.0, w_mesh=0.001)
    lp = out["loss_parts"]
    assert abs(lp[2] - float(g["mesh_loss"])) < TOL * abs(float(g["mesh_loss"]))
    want = 60.0 * float(g["kp_loss"]) + 0.001 * float(g["mesh_loss"])
    assert abs(lp[3] - want) < TOL * abs(want)
    assert rel_err(out["d_beta"], g["step_d_beta"]) < 5 * TOL      # NN index near-ties
    assert rel_err(out["d_theta"], g["step_d_theta"]) < 5 * TOL
    assert rel_err(out["d_cam"], g["step_d_cam"]) < 5 * TOL


# ------------------------------------------------------------------ stand-alone ops vs oracle
def test_batch_lbs_functions(small_model):
    rng = np.random.default_rng(11)
    th = (rng.normal(size=(50, 3)) * 0.8).astype(np.float32)
    th[0] = 0
    th[1] = [1e-9, -1e-9, 0]
    th[2] = [3.0, -2.5, 4.0]
    R = batch_lbs.batch_rodrigues(th)
    assert rel_err(R, onp.batch_rodrigues(th.astype(np.float64))) < 1e-5
    assert np.array_equal(R[0], np.eye(3, dtype=np.float32))
    assert np.array_equal(batch_lbs.batch_skew(th), onp.batch_skew(th))       # pure data movement: bit-exact
    theta = (rng.normal(size=(7, 72)) * 0.4).astype(np.float32)
    assert rel_err(batch_lbs.batch_lrotmin(theta), onp.batch_lrotmin(theta.astype(np.float64))) < 1e-5
    Rs = onp.batch_rodrigues(theta.reshape(-1, 3).astype(np.float64)).reshape(7, 24, 3, 3)
    Js = rng.normal(size=(7, 24, 3)) * 0.3
    parents = small_model["kintree_table"][0].astype(np.int32)
    newJ, A = batch_lbs.batch_global_rigid_transformation(Rs.astype(np.float32), Js.astype(np.float32), parents)
    newJ_o, A_o = onp.batch_global_rigid_transformation(Rs.astype(np.float32).astype(np.float64),
                                                       Js.astype(np.float32).astype(np.float64), parents)
    assert A.shape == (7, 24, 4, 4)
    assert rel_err(newJ, newJ_o) < 1e-5 and rel_err(A, A_o) < 1e-5
    assert np.array_equal(A[:, :, 3, :], np.broadcast_to(np.float32([0, 0, 0, 1]), (7, 24, 4)))


def test_projection_functions():
    rng = np.random.default_rng(12)
    X = rng.normal(size=(5, 37, 3)).astype(np.float32)
    cam = np.stack([rng.uniform(0.5, 1.2, 5), rng.normal(size=5) * 0.1, rng.normal(size=5) * 0.1], 1).astype(np.float32)
    p = projection.batch_orth_proj_idrot(X, cam)
    assert rel_err(p, onp.batch_orth_proj_idrot(X.astype(np.float64), cam.astype(np.float64))) < 1e-6
    q = projection.reproject_vertices(X, cam, [224.0, 224.0])
    assert rel_err(q, onp.reproject_vertices(X.astype(np.float64), cam.astype(np.float64), [224.0, 224.0])) < 1e-6
    # linear in the scale s (SURVEY §4 invariant)
    cam2 = cam.copy()
    cam2[:, 0] *= 2
    assert rel_err(projection.batch_orth_proj_idrot(X, cam2), 2 * p.astype(np.float64)) < 1e-6
    d = rng.normal(size=(5, 37, 2)).astype(np.float32)
    for im in (None, [224.0, 224.0]):
        dX, dc = projection.projection_backward(X, cam, d, im)
        if im is None:
            dX_o, dc_o = onp.orth_proj_backward(X.astype(np.float64), cam.astype(np.float64), d.astype(np.float64))
        else:
            dX_o, dc_o = onp.reproject_vertices_backward(X.astype(np.float64), cam.astype(np.float64), im, d.astype(np.float64))
        assert rel_err(dX, dX_o) < 1e-5 and rel_err(dc, dc_o) < 1e-5


def test_kp_loss_masking_bit_exact():
    inp = synthetic.make_inputs(33, seed=5)
    rng = np.random.default_rng(6)
    pred = rng.uniform(-1, 1, size=(33, 19, 2)).astype(np.float32)
    s, n, d = ops.kp_reprojection_loss_parts(inp["kp_gt"], pred, want_grad=True)
    vis = inp["kp_gt"][:, :, 2]
    assert n == 2 * int(np.count_nonzero(vis))                     # integer count: exact
    assert not d[vis == 0].any()                                   # masked terms contribute exactly 0
    assert np.array_equal(d[vis != 0], np.sign(pred - inp["kp_gt"][:, :, :2])[vis != 0])
    num, cnt = onp.kp_loss_parts(inp["kp_gt"].astype(np.float64), pred.astype(np.float64))
    assert cnt == n and abs(s - num) < 1e-5 * num
    assert abs(ops.kp_reprojection_loss(inp["kp_gt"], pred) - num / cnt) < 1e-5 * num / cnt
    # nothing visible -> exactly 0 (div_no_nan)
    gt0 = inp["kp_gt"].copy()
    gt0[:, :, 2] = 0
    assert ops.kp_reprojection_loss(gt0, pred) == 0.0
    # the weight is the visibility VALUE, the count is vis != 0 (appendix A.7)
    gt2 = np.zeros((1, 19, 3), np.float32)
    gt2[0, 0] = [0.5, 0.5, 2.0]
    assert ops.kp_reprojection_loss(gt2, np.zeros((1, 19, 2), np.float32)) == pytest.approx(1.0, rel=1e-6)


def test_mesh_loss_vs_oracle():
    rng = np.random.default_rng(21)
    B, V = 4, 700
    seg = synthetic.make_silhouettes(B, seed=9, a_range=(6, 12), b_range=(10, 20))
    pts3 = synthetic.silhouette_points(seg)
    sp = (rng.normal(size=(B, V, 2)) * np.array([12.0, 22.0]) + 112.0).astype(np.float32)
    loss, g = ops.mesh_reprojection_loss(pts3, sp, B, want_grad=True)
    lo = onp.mesh_reprojection_loss(pts3.astype(np.float64), sp.astype(np.float64), B)
    assert abs(loss - lo) < TOL * lo
    go = onp.mesh_loss_backward(pts3.astype(np.float64), sp.astype(np.float64), B)
    # per-vertex gradient parity, excluding vertices whose NN choice is a near-tie
    bad = np.abs(g - go).max(axis=2) > 1e-4 * np.abs(go).max()
    assert bad.mean() < 0.01
    assert not g[2].any()                                           # empty image: exactly 0
    # single-image helpers (ops.py:60-102)
    A = np.stack([pts3[pts3[:, 0] == 0][:, 2], pts3[pts3[:, 0] == 0][:, 1]], 1)
    iab, iba = ops.find_nearest_neighbors(A, sp[0])
    oab, oba = onp.find_nearest_neighbors(A.astype(np.float32), sp[0])
    assert (iab == oab).mean() > 0.99 and (iba == oba).mean() > 0.99
    bd = ops.bidirectional_dist(A, sp[0])
    assert abs(bd - onp.bidirectional_dist(A.astype(np.float64), sp[0].astype(np.float64))) < TOL * bd
    # exact-tie rule: duplicated vertices -> the FIRST index wins (tf.argmin)
    spd = np.concatenate([sp[0, :5], sp[0, :5]], 0)
    iab, _ = ops.find_nearest_neighbors(A, spd)
    assert iab.max() < 5


def test_gradient_penalty(golden_small):
    g = golden_small
    gin = [g["gp_in%d" % i] for i in range(4)]
    pen, sums = ops.compute_gradient_penalty(gin, want_sums=True)
    assert abs(pen - float(g["gp_penalty"])) < TOL * float(g["gp_penalty"])
    M = gin[0].shape[0]
    want = np.concatenate([x.reshape(M, -1).astype(np.float64).sum(0) for x in gin])
    assert rel_err(sums, want) < 1e-5
    assert abs(ops.gradient_penalty_from_sums(sums, M) - pen) < 1e-6 * pen
    for i, d in enumerate(ops.gradient_penalty_backward(sums, M)):
        assert d.shape == gin[i].shape
        assert rel_err(d, g["gp_grad%d" % i]) < TOL
    # sharded == whole: the sums of two halves add up (what the NCCL all-reduce does)
    h = M // 2
    _, s1 = ops.compute_gradient_penalty([x[:h] for x in gin], want_sums=True)
    _, s2 = ops.compute_gradient_penalty([x[h:] for x in gin], want_sums=True)
    assert abs(ops.gradient_penalty_from_sums(s1 + s2, M) - pen) < 1e-5 * pen


# ------------------------------------------------------------------ fresh inputs vs oracle, edge sizes
@pytest.mark.parametrize("B", [1, 5, 33, 129])
def test_step_fresh_inputs_vs_oracle(B, smpl_full, full_model):
    inp = synthetic.make_inputs(B, seed=500 + B)
    out = smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], w_kp=1.0)
    o = onp.SMPL(full_model, dtype=np.float64)
    b = {k: v.astype(np.float64) for k, v in inp.items()}
    verts, joints, Rs = o(b["beta"], b["theta"], get_skin=True)
    kp = onp.batch_orth_proj_idrot(joints, b["cam"])
    assert rel_err(out["verts"], verts) < TOL and rel_err(out["joints"], joints) < TOL and rel_err(out["Rs"], Rs) < TOL
    num, cnt = onp.kp_loss_parts(b["kp_gt"], kp)
    assert int(out["loss_parts"][1]) == cnt
    if cnt == 0:
        assert out["loss_parts"][3] == 0.0 and not out["d_theta"].any()
        return
    assert abs(out["loss_parts"][3] - num / cnt) < TOL * num / cnt
    dj, dcam = onp.orth_proj_backward(joints, b["cam"], onp.kp_loss_backward(b["kp_gt"], kp))
    db, dth = onp.smpl_backward(o, b["beta"], b["theta"], None, dj, None)
    assert rel_err(out["d_beta"], db) < TOL and rel_err(out["d_theta"], dth) < TOL and rel_err(out["d_cam"], dcam) < TOL


def test_rank1_theta_and_batch1(smpl_full, full_model):
    """data_loader.py:141 calls SMPL with beta [1,10] and theta [72]."""
    inp = synthetic.make_inputs(3, seed=77)
    verts, joints, Rs = smpl_full(inp["beta"][2:3], inp["theta"][2], get_skin=True)
    o = onp.SMPL(full_model, dtype=np.float64)
    v, j, R = o(inp["beta"][2:3].astype(np.float64), inp["theta"][2:3].astype(np.float64), get_skin=True)
    assert verts.shape == (1, 6890, 3) and rel_err(verts, v) < TOL and rel_err(joints, j) < TOL


def test_device_arrays_match_host_path_bit_exact(smpl_full):
    inp = synthetic.make_inputs(16, seed=88)
    ctx = smpl_full.ctx
    host = smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
    dev = smpl_full.step(ctx.to_device(inp["beta"]), ctx.to_device(inp["theta"]), ctx.to_device(inp["cam"]),
                         ctx.to_device(inp["kp_gt"]))
    for k in ("verts", "joints", "Rs", "kp_pred", "loss_parts", "d_beta", "d_theta", "d_cam"):
        assert isinstance(dev[k], runtime.DeviceArray)
        assert np.array_equal(dev[k].numpy(), host[k]), k
    with pytest.raises(TypeError):
        smpl_full(ctx.to_device(inp["beta"]), inp["theta"])          # mixing kinds is an error


def test_backward_requires_matching_forward(smpl_small, golden_small):
    from hpe_b200 import SmplbError
    smpl_small(golden_small["beta"], golden_small["theta"])
    with pytest.raises(SmplbError) as ei:
        smpl_small.backward(d_joints=np.zeros((2, 19, 3), np.float32))
    assert ei.value.code == -4


# ------------------------------------------------------------------ BASELINE sizes: properties
def test_full_size_properties(full_model):
    """B = 4096 (BASELINE config 2): determinism, shard-equivalence and
    identity-pose invariants, which need no oracle run at that size."""
    B = 4096
    s = SMPL(full_model, max_batch=B)
    inp = synthetic.make_inputs(B, seed=1000)
    inp["theta"][7] = 0.0
    a = s.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
    a = {k: np.array(v) for k, v in a.items()}
    b = s.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
    for k in a:
        assert np.array_equal(a[k], b[k]), "run-to-run nondeterminism in %s" % k
    assert np.isfinite(a["verts"]).all() and np.isfinite(a["d_theta"]).all()
    # per-sample outputs do not depend on the batch they are computed in; the loss numerators
    # and counts of two shards add up to the whole (the multi-GPU reduction)
    h = B // 2
    tot = int(a["loss_parts"][1])
    lo = {k: np.array(v) for k, v in s.step(inp["beta"][:h], inp["theta"][:h], inp["cam"][:h], inp["kp_gt"][:h],
                                            kp_count_override=tot).items()}
    hi = {k: np.array(v) for k, v in s.step(inp["beta"][h:], inp["theta"][h:], inp["cam"][h:], inp["kp_gt"][h:],
                                            kp_count_override=tot).items()}
    for k in ("verts", "joints", "Rs", "kp_pred", "d_beta", "d_theta", "d_cam"):
        assert np.array_equal(np.concatenate([lo[k], hi[k]]), a[k]), k
    assert int(lo["loss_parts"][1]) + int(hi["loss_parts"][1]) == tot
    assert abs(lo["loss_parts"][0] + hi["loss_parts"][0] - a["loss_parts"][0]) < 1e-5 * a["loss_parts"][0]
    # identity pose: Rs == I exactly and verts == (sum_j W_vj) * v_shaped up to fp32 rounding
    assert np.array_equal(a["Rs"][7], np.broadcast_to(np.eye(3, dtype=np.float32), (24, 3, 3)))
    o = onp.SMPL(full_model, dtype=np.float64)
    v_shaped = (inp["beta"][7].astype(np.float64) @ o.shapedirs).reshape(-1, 3) + o.v_template
    assert rel_err(a["verts"][7], v_shaped) < 1e-5
    # spot-check 3 samples of the big batch against the oracle
    idx = [0, 1234, 4095]
    v, j, R = o(inp["beta"][idx].astype(np.float64), inp["theta"][idx].astype(np.float64), get_skin=True)
    assert rel_err(a["verts"][idx], v) < TOL and rel_err(a["joints"][idx], j) < TOL


def test_tcgen05_blend_matches_fp32_gemm(smpl_full, full_model):
    """The tcgen05 fp16-split blend GEMM against the FP32 CUDA-core GEMM of the same
    contraction, and its a-priori error bound: |dv| <= u (2+eps) sum_k |pf_k||P_k|
    with u = 2^-11 for the pose term; the shape/template terms are split to fp32 accuracy."""
    inp = synthetic.make_inputs(200, seed=4242)
    ctx = smpl_full.ctx
    try:
        ctx.debug_set("blend_tc", 0)
        v0, _, _ = smpl_full(inp["beta"], inp["theta"], get_skin=True)
    finally:
        ctx.debug_set("blend_tc", 1)
    v1, _, _ = smpl_full(inp["beta"], inp["theta"], get_skin=True)
    assert np.isfinite(v1).all()
    o = onp.SMPL(full_model, dtype=np.float64)
    inter = {}
    o(inp["beta"][:8].astype(np.float64), inp["theta"][:8].astype(np.float64), get_skin=True, intermediates=inter)
    bound = 2.0 ** -11 * 2.01 * (np.abs(inter["pose_feature"]) @ np.abs(o.posedirs)).reshape(8, -1, 3)
    # skinning is a convex combination of rigid transforms: it does not amplify |dv_posed|
    assert np.abs(v1[:8].astype(np.float64) - v0[:8]).max() <= bound.max() * 1.8 + 2e-6
    assert rel_err(v1, v0) < 3e-5


def test_compact_backward_equals_dense_walk(smpl_full):
    """With no upstream d_verts the backward walks only the vertices the keypoint
    regressor touches; the skipped terms are exact zeros, so the result must match the
    dense walk to fp32 summation-order noise."""
    inp = synthetic.make_inputs(40, seed=99)
    ctx = smpl_full.ctx
    a = smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
    try:
        ctx.debug_set("compact_bwd", 0)
        b = smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
    finally:
        ctx.debug_set("compact_bwd", 1)
    for k in ("d_beta", "d_theta", "d_cam"):
        assert rel_err(a[k], b[k]) < 2e-6, k


def test_step_with_dense_mask_equals_step_with_point_lists(smpl_full):
    """smplb_step_seg: the dense mask seg [B,H,W,1] of src/trainer.py:443 compacted on the device inside the step ==
    the step fed with the host-side where(seg > 0) point lists, bit for bit (same points, same order)."""
    B = 5
    inp = synthetic.make_inputs(B, seed=71)
    seg = synthetic.make_silhouettes(B, seed=72, a_range=(6, 10), b_range=(9, 14))
    pts3 = synthetic.silhouette_points(seg)
    a = smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], silhouette=ops.silhouette_csr(pts3, B))
    a = {k: np.array(v) for k, v in a.items()}
    b = smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], seg=seg)
    ctx = smpl_full.ctx
    d = smpl_full.step(ctx.to_device(inp["beta"]), ctx.to_device(inp["theta"]), ctx.to_device(inp["cam"]),
                       ctx.to_device(inp["kp_gt"]), seg=ctx.to_device(seg.reshape(B, 224, 224)))
    for k in ("verts", "loss_parts", "d_beta", "d_theta", "d_cam"):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], d[k].numpy()), k
    assert a["loss_parts"][2] > 0


def test_tcgen05_blend_transpose_gemm_matches_fp32_gemm(smpl_full, full_model):
    """Dense backward (an upstream gradient on every vertex, src/trainer.py:502 with the mesh loss on):
    d pose_feature / d beta = dp . [posedirs | shapedirs]^T over K = 20670 as a bf16 split-precision tcgen05 GEMM
    (dp and the basis as hi + lo) against the FP32 CUDA-core GEMM of the same contraction and the fp64 oracle.
    Batch sizes straddle the 128-row M tile; the gradients span 8 orders of magnitude between samples."""
    ctx = smpl_full.ctx
    o = onp.SMPL(full_model, dtype=np.float64)
    V = full_model["v_template"].shape[0]
    for B in (3, 130):
        inp = synthetic.make_inputs(B, seed=900 + B)
        rng = np.random.default_rng(B)
        d_verts = (rng.standard_normal((B, V, 3)) * 10.0 ** rng.uniform(-6, 2, size=(B, 1, 1))).astype(np.float32)
        d_joints = rng.standard_normal((B, 19, 3)).astype(np.float32)
        smpl_full(inp["beta"], inp["theta"], get_skin=True)
        g1 = smpl_full.backward(d_verts=d_verts, d_joints=d_joints)
        try:
            ctx.debug_set("blend_bwd_tc", 0)
            smpl_full(inp["beta"], inp["theta"], get_skin=True)
            g0 = smpl_full.backward(d_verts=d_verts, d_joints=d_joints)
        finally:
            ctx.debug_set("blend_bwd_tc", 1)
        n = min(B, 6)
        b64 = {k: v[:n].astype(np.float64) for k, v in inp.items()}
        db, dth = onp.smpl_backward(o, b64["beta"], b64["theta"], d_verts[:n].astype(np.float64), d_joints[:n].astype(np.float64), None)
        for a, b, w in zip(g1, g0, (db, dth)):
            # per sample: the scales differ by orders of magnitude between samples
            for i in range(B):
                assert rel_err(a[i], b[i]) < 2e-5, (B, i)
            for i in range(n):
                assert rel_err(a[i], w[i]) < TOL, (B, i)


def test_tcgen05_skinning_matches_fp32_kernel(smpl_full):
    """T = W.A on the tensor cores (fp16 operands split into hi/lo so the product keeps
    fp32-grade accuracy) against the FP32 CUDA-core skinning kernel."""
    inp = synthetic.make_inputs(77, seed=31337)
    ctx = smpl_full.ctx
    try:
        ctx.debug_set("skin_tc", 0)
        v0, j0, _ = smpl_full(inp["beta"], inp["theta"], get_skin=True)
    finally:
        ctx.debug_set("skin_tc", 1)
    v1, j1, _ = smpl_full(inp["beta"], inp["theta"], get_skin=True)
    assert np.isfinite(v1).all()
    assert rel_err(v1, v0) < 5e-6 and rel_err(j1, j0) < 5e-6


def test_tcgen05_skinning_backward_matches_fp32_kernel(smpl_full, full_model):
    """k_skin_bwd_tc (T = W.A and dA = W^T.(g x [v_posed; 1]) as tcgen05 contractions, bf16 hi / lo operands for the
    gradient-scaled one) against k_skin_bwd on the FP32 pipes: dense backward with upstream d_verts of mesh-loss
    magnitude (1e-4: fp16 would underflow), d_joints and d_Rs; ragged batch (chunks of 8) and the ragged last vertex
    tile; and against the fp64 oracle."""
    ctx = smpl_full.ctx
    V = full_model["v_template"].shape[0]
    o = onp.SMPL(full_model, dtype=np.float64)
    for B, scale in ((5, 1.0), (61, 1e-4)):
        inp = synthetic.make_inputs(B, seed=900 + B)
        rng = np.random.default_rng(B)
        ups = dict(d_verts=(rng.standard_normal((B, V, 3)) * scale).astype(np.float32),
                   d_joints=(rng.standard_normal((B, 19, 3)) * scale).astype(np.float32),
                   d_Rs=(rng.standard_normal((B, 24, 3, 3)) * scale).astype(np.float32))
        res = {}
        try:
            for mode in (1, 0):
                ctx.debug_set("skin_bwd_tc", mode)
                smpl_full(inp["beta"], inp["theta"], get_skin=True)
                res[mode] = [np.array(x) for x in smpl_full.backward(**ups)]
        finally:
            ctx.debug_set("skin_bwd_tc", 1)
        for a, b in zip(res[1], res[0]):
            assert np.isfinite(a).all()
            assert rel_err(a, b) < 3e-5, (B, scale)
        n = min(B, 3)
        ref = onp.smpl_backward(o, inp["beta"][:n].astype(np.float64), inp["theta"][:n].astype(np.float64),
                                **{k: v[:n].astype(np.float64) for k, v in ups.items()})
        for a, b in zip(res[1], ref):
            assert rel_err(a[:n], b) < TOL


def test_fused_blend_skinning_matches_two_kernel_path(smpl_full, full_model):
    """k_body_res / k_body_pair / k_body_tc (blend + skinning in one kernel, v_posed stays in TMEM) against k_blend_tc +
    k_skin_tc: the same fp16 operands and fp32 accumulation, so the two may differ only by fp32
    summation order; and against the fp64 oracle at the stated tolerance.  Batch sizes straddle
    the 96-sample super-tile and the 8-sample skinning tile (ragged tails)."""
    ctx = smpl_full.ctx
    o = onp.SMPL(full_model, dtype=np.float64)
    for B in (1, 7, 96, 203, 1700):     # 1700: 18 sample blocks x 27 vertex-tile pairs, pairs straddle vertex tiles
        inp = synthetic.make_inputs(B, seed=500 + B)
        try:
            ctx.debug_set("fused", 0)
            v0, j0, _ = smpl_full(inp["beta"], inp["theta"], get_skin=True)
            # 1: the default (CTA pairs, Dt16 tile resident in shared memory, W16 in tensor memory, k_body_res); 10: the
            # same with shallower operand rings; 9: sixteen epilogue warps;
            # 7: CTA pairs streaming Dt16 (k_body_pair); 4: the best single-CTA configuration; 2, 3: tuning
            # variants; 5: W16 as a TMEM-resident operand; 6: CTA pairs with the Dt16 tiles multicast
            v_tma = None
            for variant in (1, 10, 9, 7, 2, 3, 4, 5, 6):
                ctx.debug_set("fused", variant)
                v1, j1, _ = smpl_full(inp["beta"], inp["theta"], get_skin=True)
                assert np.isfinite(v1).all()
                # 1 and 10 differ only in the depth of the operand rings: same MMAs in the same order, same bits
                if variant == 1:
                    v_tma = v1
                if variant == 10:
                    assert np.array_equal(v1, v_tma), (B, "ring depth changed the result")
                assert rel_err(v1, v0) < 2e-6, (B, variant)
                assert np.array_equal(j1, j0)
                n = min(B, 4)
                v, _, _ = o(inp["beta"][:n].astype(np.float64), inp["theta"][:n].astype(np.float64), get_skin=True)
                assert rel_err(v1[:n], v) < TOL
        finally:
            ctx.debug_set("fused", 1)
    # the dense backward after a fused forward rebuilds v_posed on demand
    inp = synthetic.make_inputs(33, seed=77)
    rng = np.random.default_rng(5)
    d_verts = rng.standard_normal((33, full_model["v_template"].shape[0], 3)).astype(np.float32)
    smpl_full(inp["beta"], inp["theta"], get_skin=True)
    g1 = smpl_full.backward(d_verts=d_verts)
    try:
        ctx.debug_set("fused", 0)
        smpl_full(inp["beta"], inp["theta"], get_skin=True)
        g0 = smpl_full.backward(d_verts=d_verts)
    finally:
        ctx.debug_set("fused", 1)
    for a, b in zip(g1, g0):
        assert rel_err(a, b) < 2e-6


def test_forward_only_sm_split_does_not_change_verts(smpl_full):
    """A forward-only call gives the vertex kernel 64 of the 74 SM pairs so that the keypoint path runs beside it
    (`pairs_auto`); how the super-tiles are dealt to the CTAs must not change a bit of the result."""
    ctx = smpl_full.ctx
    inp = synthetic.make_inputs(64, seed=321)
    res = {}
    try:
        for pairs in (0, -1, 17):        # automatic (64), every SM pair, an odd count that straddles vertex tiles
            ctx.debug_set("body_pairs", pairs)
            res[pairs] = smpl_full(inp["beta"], inp["theta"], get_skin=True)
    finally:
        ctx.debug_set("body_pairs", 0)
    for pairs in (-1, 17):
        for a, b in zip(res[0], res[pairs]):
            assert np.array_equal(a, b), pairs


def test_fused_blend_skinning_small_vertex_counts():
    """The CTA-pair kernel needs an even number of 128-vertex tiles: 200 vertices = 2 tiles = one
    pair (its second tile mostly padding), 300 = 3 tiles -> single-CTA fallback, 700 = 6 tiles = 3
    pairs; each against the two-kernel path and the fp64 oracle."""
    for nv in (200, 300, 700):
        model = synthetic.make_model(seed=nv, num_verts=nv, regressor_nnz=8)
        s = SMPL(model, max_batch=128)
        o = onp.SMPL(model, dtype=np.float64)
        for B in (3, 101):
            inp = synthetic.make_inputs(B, seed=nv + B)
            s.ctx.debug_set("fused", 0)
            v0, j0, _ = s(inp["beta"], inp["theta"], get_skin=True)
            s.ctx.debug_set("fused", 1)
            v1, j1, _ = s(inp["beta"], inp["theta"], get_skin=True)
            assert v1.shape == (B, nv, 3) and np.isfinite(v1).all()
            assert rel_err(v1, v0) < 2e-6, (nv, B)
            v, _, _ = o(inp["beta"][:3].astype(np.float64), inp["theta"][:3].astype(np.float64), get_skin=True)
            assert rel_err(v1[:3], v) < TOL


def test_pose_backward_subtree_sum_form_matches_serial_walk(smpl_full, full_model):
    """k_pose_bwd_reg (the reverse kinematic chain as subtree sums in registers, the default) against k_pose_bwd (the
    23-step walk of batch_lbs.py:128-135 reversed joint by joint): the keypoint step, and the dense backward with
    upstream d_verts / d_joints / d_Rs -- same inputs, different summation order only."""
    ctx = smpl_full.ctx
    inp = synthetic.make_inputs(61, seed=4242)
    rng = np.random.default_rng(9)
    V = full_model["v_template"].shape[0]
    ups = dict(d_verts=rng.standard_normal((61, V, 3)).astype(np.float32) * 1e-3,
               d_joints=rng.standard_normal((61, 19, 3)).astype(np.float32),
               d_Rs=rng.standard_normal((61, 24, 3, 3)).astype(np.float32))
    res = {}
    try:
        for mode in (1, 0):
            ctx.debug_set("pose_bwd_reg", mode)
            st = smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
            smpl_full(inp["beta"], inp["theta"], get_skin=True)
            g = smpl_full.backward(**ups)
            res[mode] = [np.array(st[k]) for k in ("d_beta", "d_theta", "d_cam")] + [np.array(x) for x in g]
    finally:
        ctx.debug_set("pose_bwd_reg", 1)
    for a, b in zip(res[1], res[0]):
        assert np.isfinite(a).all()
        assert rel_err(a, b) < 5e-6


def test_fused_keypoint_forward_backward_matches_separate_kernels(smpl_full):
    """k_fold_step_w (forward + backward of the folded keypoint path in one kernel, gradients formed
    for a unit loss scale and scaled by w_kp / num_present in k_pose_bwd) against the separate
    kernels, including the all-invisible sample, a count override (the multi-GPU shard case) and a
    non-default loss weight."""
    inp = synthetic.make_inputs(150, seed=2024)
    ctx = smpl_full.ctx
    for kw in ({}, {"kp_count_override": 12345}, {"w_kp": 7.5}):
        a = {k: np.array(v) for k, v in smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], **kw).items()}
        try:
            ctx.debug_set("fold_step", 0)
            b = {k: np.array(v) for k, v in smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], **kw).items()}
        finally:
            ctx.debug_set("fold_step", 1)
        for k in ("joints", "kp_pred", "verts", "Rs"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a["loss_parts"][:2], b["loss_parts"][:2])
        assert abs(a["loss_parts"][3] - b["loss_parts"][3]) <= 1e-6 * abs(b["loss_parts"][3])
        for k in ("d_beta", "d_theta", "d_cam"):
            assert rel_err(a[k], b[k]) < 2e-6, (k, kw)
        assert not a["d_theta"][1].any() and not a["d_cam"][1].any()     # sample 1 has no visible keypoint


def test_many_keypoints_and_trace_mode():
    """(a) A regressor with more keypoints than k_fold_step_w stages in shared memory (24 * 3 * K
    floats > 1536, i.e. K > 21) takes the separate forward / backward kernels: results against the
    oracle.  (b) The timeline trace (smplb_profile_enable(ctx, 2)) lists every launch of a step
    with start <= end, without serialising the streams."""
    import scipy.sparse as sp
    model = synthetic.make_model(seed=3, num_verts=300, regressor_nnz=8)
    rng = np.random.default_rng(9)
    extra = synthetic._sparse_regressor(rng, 7, 300, 8)
    model["cocoplus_regressor"] = sp.vstack([model["cocoplus_regressor"], extra]).tocsc()
    s = SMPL(model, max_batch=16)
    K = s.num_keypoints
    assert K == 26
    B = 11
    inp = synthetic.make_inputs(B, seed=5)
    kp_gt = np.concatenate([inp["kp_gt"], inp["kp_gt"][:, :7]], axis=1).astype(np.float32)
    s.ctx.profile(2)
    out = s.step(inp["beta"], inp["theta"], inp["cam"], kp_gt, w_kp=1.0)
    trace = s.ctx.profile_trace()
    s.ctx.profile(0)
    names = [n for n, _, _ in trace]
    assert "pose_fwd" in names and "pose_bwd" in names and "fold_step_fwd_bwd" not in names
    assert all(t1 >= t0 for _, t0, t1 in trace)
    o = onp.SMPL(model, dtype=np.float64)
    b = {k: v.astype(np.float64) for k, v in inp.items()}
    verts, joints, Rs = o(b["beta"], b["theta"], get_skin=True)
    kp = onp.batch_orth_proj_idrot(joints, b["cam"])
    assert rel_err(out["verts"], verts) < TOL and rel_err(out["joints"], joints) < TOL
    num, cnt = onp.kp_loss_parts(kp_gt.astype(np.float64), kp)
    assert int(out["loss_parts"][1]) == cnt
    assert abs(out["loss_parts"][3] - num / cnt) < TOL * num / cnt
    dj, dcam = onp.orth_proj_backward(joints, b["cam"], onp.kp_loss_backward(kp_gt.astype(np.float64), kp))
    db, dth = onp.smpl_backward(o, b["beta"], b["theta"], None, dj, None)
    assert rel_err(out["d_beta"], db) < TOL and rel_err(out["d_theta"], dth) < TOL and rel_err(out["d_cam"], dcam) < TOL


def test_folded_keypoint_path_matches_per_vertex_path(smpl_full):
    """joints / kp loss / gradients from the folded formulation (G x, no vertices) against the
    per-vertex keypoint path; both are checked against the oracle elsewhere, this pins them to
    each other at a tighter level."""
    inp = synthetic.make_inputs(150, seed=2718)
    ctx = smpl_full.ctx
    a = smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
    a = {k: np.array(v) for k, v in a.items()}
    try:
        ctx.debug_set("fold", 0)
        b = smpl_full.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
    finally:
        ctx.debug_set("fold", 1)
    assert np.array_equal(a["verts"], b["verts"])                    # the verts path is untouched
    # (the pose-feature columns of both formulations are single fp16 products, bound 2^-11 * 2 * sum |G||pf|)
    for k, tol in (("joints", 3e-5), ("kp_pred", 3e-5), ("d_beta", 3e-5), ("d_theta", 3e-5), ("d_cam", 3e-5)):
        assert rel_err(a[k], b[k]) < tol, k
    assert a["loss_parts"][1] == b["loss_parts"][1]
    assert abs(a["loss_parts"][3] - b["loss_parts"][3]) < 1e-5 * abs(b["loss_parts"][3])
    # joints-only forward (get_skin=False) never touches the 6890-vertex tensors
    j = smpl_full(inp["beta"], inp["theta"])
    assert rel_err(j, a["joints"]) < 1e-6
    assert not smpl_full.ctx.profile_read()                          # (profiling is off: nothing recorded)
    # backward with upstream on joints and Rs only goes through the folded path
    rng = np.random.default_rng(3)
    dj = rng.normal(size=(150, 19, 3)).astype(np.float32)
    dR = rng.normal(size=(150, 24, 3, 3)).astype(np.float32)
    smpl_full(inp["beta"], inp["theta"], get_skin=True)
    db1, dt1 = smpl_full.backward(d_joints=dj, d_Rs=dR)
    try:
        ctx.debug_set("fold", 0)
        smpl_full(inp["beta"], inp["theta"], get_skin=True)
        db0, dt0 = smpl_full.backward(d_joints=dj, d_Rs=dR)
    finally:
        ctx.debug_set("fold", 1)
    assert rel_err(db1, db0) < 2e-5 and rel_err(dt1, dt0) < 2e-5


def test_mesh_grid_search_equals_brute_force():
    """The accelerated nearest-neighbour searches must reproduce the reference's full scan bit for bit: same indices
    (first index on ties), same loss, same gradient.  Modes: lattice (vertex -> pixel on the pixel bitmap, the default
    for row-major pixel lists) + binned grid, binned grid only, brute force.  Image 4's points are shuffled and image
    5's are not integers, so those two fall back to the grid inside the lattice mode."""
    rng = np.random.default_rng(77)
    B, V = 6, 1500
    seg = synthetic.make_silhouettes(B, seed=19, a_range=(10, 30), b_range=(20, 60))
    pts3 = synthetic.silhouette_points(seg)
    sel4 = np.flatnonzero(pts3[:, 0] == 4)
    pts3[sel4] = pts3[rng.permutation(sel4)]           # not in row-major order: no lattice for this image
    pts3[pts3[:, 0] == 5, 1:] += 0.25                  # not on the integer lattice
    sp = (rng.normal(size=(B, V, 2)) * np.array([25.0, 45.0]) + 112.0).astype(np.float32)
    sp[0, :200] = sp[0, 200:400]                       # exact duplicates: ties must pick the first index
    sp[0, 400:500] = np.round(sp[0, 400:500])          # vertices exactly on lattice points: equidistant pixels
    sp[0, 500:600] = np.round(sp[0, 500:600]) + 0.5    # ... and exactly between them
    sp[1] += 400.0                                     # a mesh projected far outside the image
    sp[3, :, 0] = 100.0                                # degenerate: all vertices on one vertical line
    ctx = ops._ctx_for(sp)
    pts, offs = ops.silhouette_csr(pts3, B)
    res = {}
    try:
        for mode in ("lattice", "grid", "brute"):
            ctx.debug_set("mesh_grid", 0 if mode == "brute" else 1)
            ctx.debug_set("mesh_lattice", 1 if mode == "lattice" else 0)
            res[mode] = ops._mesh_call(ctx, pts, offs, sp, True, True)
    finally:
        ctx.debug_set("mesh_grid", 1)
        ctx.debug_set("mesh_lattice", 1)
    l0, g0, ab0, ba0 = res["brute"]
    for mode in ("lattice", "grid"):
        l1, g1, ab1, ba1 = res[mode]
        assert np.array_equal(ab1, ab0) and np.array_equal(ba1, ba0), mode
        assert l1[0] == l0[0], mode
        assert np.array_equal(g1, g0), mode
    # and all agree with the oracle's value
    lo = onp.mesh_reprojection_loss(pts3.astype(np.float64), sp.astype(np.float64), B)
    assert abs(l0[0] - lo) < TOL * lo


def test_section_8f_rows():
    """Neighbours of the path: on-device silhouette extraction, KCS forward/backward,
    critic-input interpolation."""
    from hpe_b200 import models
    B = 5
    seg = synthetic.make_silhouettes(B, seed=23, a_range=(8, 20), b_range=(12, 40))
    pts_h, offs_h = ops.silhouette_csr(synthetic.silhouette_points(seg), B)
    pts_d, offs_d = ops.silhouette_csr_device(seg)
    assert np.array_equal(offs_d, offs_h)                           # integer work: bit-exact
    assert np.array_equal(pts_d[:offs_h[-1]], pts_h)                # same points, same (row-major) order
    pts_c, offs_c = ops.silhouette_csr_device(seg, cap=100)         # capacity-limited: true counts, truncated list
    assert np.array_equal(offs_c, offs_h) and np.array_equal(pts_c, pts_h[:100])
    rng = np.random.default_rng(5)
    joints = rng.normal(size=(9, 19, 3)).astype(np.float32)
    C = models.precompute_C_matrix()
    assert np.array_equal(C, onp.precompute_C_matrix().astype(np.float32))
    k = models.get_kcs(joints, C)
    assert k.shape == (9, 13, 13) and rel_err(k, onp.get_kcs(joints.astype(np.float64), C.astype(np.float64))) < 1e-5
    up = rng.normal(size=(9, 13, 13)).astype(np.float32)
    dj = models.get_kcs_backward(joints, C, up)
    dj_o = onp.get_kcs_backward(joints.astype(np.float64), C.astype(np.float64), up.astype(np.float64))
    assert rel_err(dj, dj_o) < 1e-5 and not dj[:, 14:].any()
    fake, real = rng.normal(size=(9, 23, 3, 3)).astype(np.float32), rng.normal(size=(9, 23, 3, 3)).astype(np.float32)
    alpha = rng.uniform(size=9).astype(np.float32)
    assert rel_err(models.interpolate(fake, real, alpha), fake + alpha[:, None, None, None] * (real - fake)) < 1e-6
    alpha_e = rng.uniform(size=fake.shape).astype(np.float32)     # element-wise, as trainer.py:548-550 draws it
    assert rel_err(models.interpolate(fake, real, alpha_e), fake + alpha_e * (real - fake)) < 1e-6


def test_fused_critic_inputs_and_gradient_penalty():
    """SURVEY section 8f rank 1, fused: (1) the interpolated critic inputs + get_kcs of the interpolated joints in one
    launch (trainer.py:548-557, models.py:123-139); (2) the gradient w.r.t. the joints completed with the path
    through get_kcs and the gradient penalty over all four gradients in one launch (trainer.py:566-572,
    ops.py:153-172) -- against the oracle and against the separate calls."""
    from hpe_b200 import models
    rng = np.random.default_rng(17)
    M = 333                                                          # 6 chunks of 64 rows, the last one ragged
    C = models.precompute_C_matrix()
    mk = lambda *shape: rng.normal(size=shape).astype(np.float32)    # noqa: E731
    un = lambda *shape: rng.uniform(size=shape).astype(np.float32)   # noqa: E731
    fj, rj, aj = mk(M, 14, 3), mk(M, 14, 3), un(M, 14, 3)
    fs, rs, as_ = mk(M, 10), mk(M, 10), un(M, 10)
    fR, rR, aR = mk(M, 23, 3, 3), mk(M, 23, 3, 3), un(M, 23, 3, 3)
    jh, kh, sh, Rh = models.critic_inputs(fj, rj, aj, fs, rs, as_, fR, rR, aR, C)
    jw = fj.astype(np.float64) + aj * (rj.astype(np.float64) - fj)
    assert rel_err(jh, jw) < 1e-6 and rel_err(sh, fs + as_ * (rs - fs)) < 1e-6 and rel_err(Rh, fR + aR * (rR - fR)) < 1e-6
    assert rel_err(kh, onp.get_kcs(jw, C.astype(np.float64))) < 1e-5
    # the critic's partial derivatives (any values: the critic itself is out of scope)
    g_kcs, g_j, g_s, g_R = mk(M, 13, 13) + 0.1, mk(M, 14, 3) + 0.1, mk(M, 10) + 0.1, mk(M, 23, 3, 3) + 0.1
    res = models.critic_gradient_penalty(jh, C, g_kcs, g_j, g_s, g_R, want=("col_sums", "g_joints_total"))
    gj_total = g_j.astype(np.float64) + onp.get_kcs_backward(jh.astype(np.float64), C.astype(np.float64), g_kcs.astype(np.float64))
    assert rel_err(res["g_joints_total"], gj_total) < 1e-5
    want = onp.compute_gradient_penalty([g_kcs.astype(np.float64), gj_total, g_s.astype(np.float64), g_R.astype(np.float64)])
    assert abs(res["penalty"] - want) < 1e-5 * want
    # == the separate calls (kcs backward, add, penalty), and repeatable bit for bit (fixed-order partial sums)
    sep = ops.compute_gradient_penalty([g_kcs, g_j + models.get_kcs_backward(jh, C, g_kcs), g_s, g_R])
    assert abs(res["penalty"] - sep) < 1e-5 * sep
    again = models.critic_gradient_penalty(jh, C, g_kcs, g_j, g_s, g_R, want=("col_sums",))
    assert again["penalty"] == res["penalty"] and np.array_equal(again["col_sums"], res["col_sums"])


def test_mocap_preprocessing_batched(smpl_full, full_model):
    """data_loader.py:139-143 maps SMPL over the mocap table one pose at a time; here the table
    is one batched call and must give what per-example calls give."""
    from hpe_b200 import data_loader
    inp = synthetic.make_inputs(300, seed=606)
    joints, shape, rots = data_loader.preprocess_poses(smpl_full, inp["theta"], inp["beta"], chunk=128)
    assert joints.shape == (300, 19, 3) and rots.shape == (300, 24, 3, 3) and shape.shape == (300, 10)
    for i in (0, 1, 131, 299):
        v, j, R = smpl_full(inp["beta"][i:i + 1], inp["theta"][i], get_skin=True)     # the reference's per-example call
        assert rel_err(joints[i], j[0]) < 1e-5 and np.array_equal(rots[i], R[0])
    o = onp.SMPL(full_model, dtype=np.float64)
    _, j64, R64 = o(inp["beta"][:4].astype(np.float64), inp["theta"][:4].astype(np.float64), get_skin=True)
    assert rel_err(joints[:4], j64) < TOL and rel_err(rots[:4], R64) < TOL


def test_steps_repeat_bit_exactly_with_contexts_in_flight(full_model):
    """A step uses four streams per context (per-body kernels, the vertex kernel, the tcgen05 GEMMs
    on a low-priority stream, the loss reduction) and several contexts share the GPU: with three
    contexts taking steps back to back over rotating inputs, every step's loss and gradients must
    equal, bit for bit, the first result for the same (context, inputs) -- a missing dependency
    between the streams shows up as a difference (tools/determinism.py is the long version, also
    for torchrun)."""
    B, NE, NSET, STEPS = 1024, 3, 3, 150
    engines = [SMPL(full_model, max_batch=B) for _ in range(NE)]
    host_sets = [synthetic.make_inputs(B, seed=4000 + i) for i in range(NSET)]
    dev_sets = [[{k: e.ctx.to_device(v) for k, v in s.items()} for s in host_sets] for e in engines]
    depth = 2 * NE
    outs = [{} for _ in range(depth)]
    ref, pending, bad = {}, [], []

    def check(item):
        key, o = item
        engines[key[0]].ctx.sync()
        got = tuple(o[n].numpy().tobytes() for n in ("loss_parts", "d_theta", "d_beta", "d_cam"))
        if ref.setdefault(key, got) != got:
            bad.append(key)

    for i in range(STEPS):
        e, s = i % NE, (i // NE) % NSET
        if len(pending) >= depth:
            check(pending.pop(0))
        d = dev_sets[e][s]
        engines[e].step(d["beta"], d["theta"], d["cam"], d["kp_gt"], w_kp=60.0, out=outs[i % depth])
        pending.append(((e, s), outs[i % depth]))
    while pending:
        check(pending.pop(0))
    assert not bad, bad[:5]
    # and the three contexts agree with each other
    for s in range(NSET):
        assert ref[(0, s)] == ref[(1, s)] == ref[(2, s)]
