"""CPU tests that pin the oracle (oracle/smpl_numpy.py):
  * against the committed golden vectors, which are outputs of the reference's
    own code (oracle/make_golden.py), and
  * live against the reference executed under the tf shim where
    /root/reference exists (build container only).
The reference itself ships no tests or fixtures."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import smpl_numpy as onp
from oracle import run_reference

TOL64 = 1e-11


def compact(pts, B):
    keep = [i for i in range(B) if np.any(pts[:, 0] == i)]
    out = []
    for n, i in enumerate(keep):
        r = pts[pts[:, 0] == i].copy()
        r[:, 0] = n
        out.append(r)
    return np.concatenate(out), keep


def check_case(model, g, full=False, stride=1):
    o = onp.SMPL(model, dtype=np.float64)
    beta, theta, cam = (g[k].astype(np.float64) for k in ("beta", "theta", "cam"))
    kp_gt = g["kp_gt"].astype(np.float64)
    B = beta.shape[0]
    verts, joints, Rs = o(beta, theta, get_skin=True)
    sub = (lambda x: x[:, ::stride]) if full else (lambda x: x)
    sfx = "_sub" if full else ""
    assert rel_err(sub(verts), g["verts" + sfx]) < TOL64
    assert rel_err(joints, g["joints"]) < TOL64
    assert rel_err(Rs, g["Rs"]) < TOL64
    assert rel_err(o.J_transformed, g["J_transformed"]) < TOL64
    kp = onp.batch_orth_proj_idrot(joints, cam)
    assert rel_err(kp, g["kp_pred"]) < TOL64
    assert abs(onp.kp_reprojection_loss(kp_gt, kp) - float(g["kp_loss"])) < TOL64
    sp = onp.reproject_vertices(verts, cam, [224.0, 224.0])
    assert rel_err(sub(sp), g["sil_pred" + sfx]) < TOL64
    pts = g["seg_points"].astype(np.float64)
    ml = onp.mesh_reprojection_loss(pts, sp, B)
    assert abs(ml - float(g["mesh_loss"])) < 1e-9 * abs(float(g["mesh_loss"]))
    # gradients
    dkp = onp.kp_loss_backward(kp_gt, kp)
    dj, dcam = onp.orth_proj_backward(joints, cam, dkp)
    db, dth = onp.smpl_backward(o, beta, theta, None, dj, None)
    assert rel_err(db, g["kp_d_beta"]) < 1e-9
    assert rel_err(dth, g["kp_d_theta"]) < 1e-9
    assert rel_err(dcam, g["kp_d_cam"]) < 1e-9
    dsp = onp.mesh_loss_backward(pts, sp, B)
    pk, keep = compact(pts, B)
    assert rel_err(sub(dsp), g["mesh_d_sil_pred" + sfx]) < 1e-9
    dv, dcam_m = onp.reproject_vertices_backward(verts, cam, [224.0, 224.0], 0.001 * dsp)
    db, dth = onp.smpl_backward(o, beta, theta, dv, 60.0 * dj, None)
    assert rel_err(db, g["step_d_beta"]) < 1e-9
    assert rel_err(dth, g["step_d_theta"]) < 1e-9
    assert rel_err(60.0 * dcam + dcam_m, g["step_d_cam"]) < 1e-9
    db, dth = onp.smpl_backward(o, beta, theta, g["up_verts"], g["up_joints"], g["up_Rs"])
    assert rel_err(db, g["up_d_beta"]) < 1e-9
    assert rel_err(dth, g["up_d_theta"]) < 1e-9
    gp_in = [g["gp_in%d" % i].astype(np.float64) for i in range(4)]
    assert abs(onp.compute_gradient_penalty(gp_in) - float(g["gp_penalty"])) < TOL64
    for i, d in enumerate(onp.gradient_penalty_backward(gp_in)):
        assert rel_err(d, g["gp_grad%d" % i]) < 1e-9


def test_oracle_matches_golden_small(small_model, golden_small):
    check_case(small_model, golden_small)


def test_oracle_matches_golden_full(full_model, golden_full):
    from oracle.make_golden import model_digest
    assert model_digest(full_model) == str(golden_full["model_sha256"]), "synthetic model recipe drifted"
    check_case(full_model, golden_full, full=True, stride=int(golden_full["vert_stride"]))


def test_oracle_lsp(small_model, golden_small_lsp):
    g = golden_small_lsp
    o = onp.SMPL(small_model, joint_type="lsp")
    joints = o(g["beta"].astype(np.float64), g["theta"].astype(np.float64))
    assert joints.shape[1] == 14
    assert rel_err(joints, g["joints"]) < TOL64
    kp = onp.batch_orth_proj_idrot(joints, g["cam"].astype(np.float64))
    assert abs(onp.kp_reprojection_loss(g["kp_gt"].astype(np.float64), kp) - float(g["kp_loss"])) < TOL64


def test_invariants(small_model):
    """SURVEY.md §4 known-answer invariants."""
    o = onp.SMPL(small_model)
    beta = np.random.default_rng(0).normal(size=(3, 10))
    theta = np.zeros((3, 72))
    inter = {}
    verts, joints, Rs = o(beta, theta, get_skin=True, intermediates=inter)
    assert np.max(np.abs(Rs - np.eye(3))) < 3e-16                        # theta = 0 -> I (cos(1.7e-8) rounds in fp64)
    R32 = onp.batch_rodrigues(np.zeros((2, 3), dtype=np.float32))
    assert np.array_equal(R32, np.broadcast_to(np.eye(3, dtype=np.float32), R32.shape))   # exactly I in fp32
    assert rel_err(verts, inter["v_posed"]) < 1e-6       # identity pose: verts = (sum_j W_vj) v_posed, fp32-rounded weights
    A = inter["A"]
    assert np.array_equal(A[:, :, 3, :], np.broadcast_to(np.array([0, 0, 0, 1.0]), A[:, :, 3, :].shape))
    # fold used by the kernels: J = J_regressor^T v_template + (J_regressor^T shapedirs) beta
    V = o.size[0]
    J0 = o.J_regressor.T @ o.v_template
    Jd = np.einsum("vj,kvc->jck", o.J_regressor, o.shapedirs.reshape(10, V, 3))
    assert rel_err(J0[None] + np.einsum("jck,bk->bjc", Jd, beta), inter["J"]) < 1e-13
    # all-invisible keypoints -> loss exactly 0 (div_no_nan)
    kp_gt = np.zeros((2, 19, 3))
    assert onp.kp_reprojection_loss(kp_gt, np.ones((2, 19, 2))) == 0.0
    # visibility weights multiply by the VALUE, count uses != 0
    kp_gt[0, 0] = [0.5, 0.5, 2.0]
    assert onp.kp_reprojection_loss(kp_gt, np.zeros((2, 19, 2))) == pytest.approx(2.0 * 1.0 / 2)


def test_fp32_oracle_close_to_fp64(small_model, golden_small):
    o = onp.SMPL(small_model, dtype=np.float32)
    verts, joints, Rs = o(golden_small["beta"], golden_small["theta"], get_skin=True)
    assert verts.dtype == np.float32
    assert rel_err(verts, golden_small["verts"]) < 5e-6
    assert rel_err(joints, golden_small["joints"]) < 5e-6


@pytest.mark.skipif(not run_reference.available(), reason="reference checkout not mounted (GPU box)")
def test_oracle_matches_live_reference(small_model):
    """Oracle A (reference files run unchanged) == Oracle B, fresh inputs."""
    import warnings
    warnings.filterwarnings("ignore")
    from hpe_b200 import synthetic
    ref = run_reference.Reference(float64=True)
    rs = ref.load_smpl(small_model)
    o = onp.SMPL(small_model)
    inp = synthetic.make_inputs(4, seed=77, dtype=np.float64)
    beta, theta, cam = ref.tensor(inp["beta"], True), ref.tensor(inp["theta"], True), ref.tensor(inp["cam"], True)
    verts, joints, Rs = rs(beta, theta, get_skin=True)
    v2, j2, R2 = o(inp["beta"], inp["theta"], get_skin=True)
    assert rel_err(v2, ref.np(verts)) < 1e-13 and rel_err(j2, ref.np(joints)) < 1e-13 and rel_err(R2, ref.np(Rs)) < 1e-13
    kp = ref.projection.batch_orth_proj_idrot(joints, cam)
    loss = ref.ops.kp_reprojection_loss(ref.tensor(inp["kp_gt"]), kp)
    gb, gt, gc = ref.torch.autograd.grad(loss, [beta, theta, cam])
    kp2 = onp.batch_orth_proj_idrot(j2, inp["cam"])
    dj, dcam = onp.orth_proj_backward(j2, inp["cam"], onp.kp_loss_backward(inp["kp_gt"], kp2))
    db, dth = onp.smpl_backward(o, inp["beta"], inp["theta"], None, dj, None)
    assert rel_err(db, gb.numpy()) < 1e-10 and rel_err(dth, gt.numpy()) < 1e-10 and rel_err(dcam, gc.numpy()) < 1e-10
    # reference-side helper functions
    th = ref.tensor(inp["theta"].reshape(-1, 3))
    assert rel_err(onp.batch_rodrigues(inp["theta"].reshape(-1, 3)), ref.np(ref.batch_lbs.batch_rodrigues(th))) < 1e-14
    assert rel_err(onp.batch_skew(inp["theta"].reshape(-1, 3)), ref.np(ref.batch_lbs.batch_skew(th))) < 1e-14
    assert rel_err(onp.batch_lrotmin(inp["theta"]), ref.np(ref.batch_lbs.batch_lrotmin(ref.tensor(inp["theta"])))) < 1e-14


@pytest.mark.skipif(not run_reference.available(), reason="reference checkout not mounted (GPU box)")
def test_kcs_oracle_matches_reference_source():
    """get_kcs / precompute_C_matrix of src/models.py executed from the reference's own source
    text (the module itself imports keras, so the two functions are lifted with ast)."""
    import ast
    import os
    import warnings
    warnings.filterwarnings("ignore")
    ref = run_reference.Reference(float64=True)
    src = open(os.path.join(run_reference.REFERENCE_ROOT, "src", "models.py")).read()
    tree = ast.parse(src)
    ns = {"tf": ref.tf, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("get_kcs", "precompute_C_matrix"):
            exec(compile(ast.Module([node], []), "models.py", "exec"), ns)
    C_ref = ref.np(ns["precompute_C_matrix"]())
    assert np.array_equal(C_ref, onp.precompute_C_matrix())
    rng = np.random.default_rng(4)
    joints = rng.normal(size=(5, 19, 3))
    jt = ref.tensor(joints, True)
    k_ref = ns["get_kcs"](jt, ref.tf.constant(C_ref, ref.dtype))
    assert rel_err(onp.get_kcs(joints, C_ref), ref.np(k_ref)) < 1e-13
    up = rng.normal(size=(5, 13, 13))
    g = ref.torch.autograd.grad((k_ref * ref.tensor(up)).sum(), [jt])[0].numpy()
    assert rel_err(onp.get_kcs_backward(joints, C_ref, up), g) < 1e-12


def test_torch_port_matches_numpy_oracle(small_model):
    """oracle/smpl_torch.py (the CPU baseline bench.py times on the GPU box) == oracle/smpl_numpy.py:
    forward, loss and autograd gradients vs the hand-derived backward, fp64."""
    import torch
    from hpe_b200 import synthetic
    from oracle import smpl_numpy as onp
    from oracle import smpl_torch as ot
    inp = synthetic.make_inputs(5, seed=11, dtype=np.float64)
    ts = ot.SMPL(small_model, dtype=torch.float64)
    verts, joints, Rs, kp, loss, db, dth, dc = ot.step(ts, inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"])
    o = onp.SMPL(small_model, dtype=np.float64)
    v, j, R = o(inp["beta"], inp["theta"], get_skin=True)
    kpo = onp.batch_orth_proj_idrot(j, inp["cam"])
    assert np.max(np.abs(verts - v)) < 1e-12 and np.max(np.abs(joints - j)) < 1e-12 and np.max(np.abs(Rs - R)) < 1e-12
    assert abs(loss - onp.kp_reprojection_loss(inp["kp_gt"], kpo)) < 1e-12
    dj, dcam = onp.orth_proj_backward(j, inp["cam"], onp.kp_loss_backward(inp["kp_gt"], kpo))
    dbo, dtho = onp.smpl_backward(o, inp["beta"], inp["theta"], None, dj, None)
    assert np.max(np.abs(db - dbo)) < 1e-10 and np.max(np.abs(dth - dtho)) < 1e-10 and np.max(np.abs(dc - dcam)) < 1e-10
