"""GPU parity at the sizes BASELINE.json names, against the fp64 oracle (not only properties):

  config 2  B = 4096: loss of the whole batch and gradients of 64 sampled bodies
  config 3  a mesh-loss step on 16 images with ~10 k silhouette pixels each and V = 6890
            (src/ops.py:117-137 through src/trainer.py:436-450): value and d_beta / d_theta / d_cam
  config 5  B = 65536 forward (src/predictor.py:141): spot check of 32 bodies
  lsp       joint_type='lsp' (K = 14, batch_smpl.py:80-81) on the 6890-vertex model

Tolerance as everywhere: max|x - ref| <= 1e-4 * max|ref| per tensor; counts bit-exact.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err
from hpe_b200 import ops, synthetic
from hpe_b200._lib import check, lib
from hpe_b200.tf_smpl import projection
from hpe_b200.tf_smpl.batch_smpl import SMPL
from oracle import smpl_numpy as onp

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _oracle_joints(o, beta, theta, chunk=128):
    out = []
    for s in range(0, beta.shape[0], chunk):
        out.append(o(beta[s:s + chunk], theta[s:s + chunk]))
    return np.concatenate(out)


def test_config2_b4096_loss_and_gradients_vs_oracle(full_model):
    B = 4096
    w_kp = 60.0
    s = SMPL(full_model, max_batch=B)
    inp = synthetic.make_inputs(B, seed=1000)
    out = s.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], w_kp=w_kp)
    o = onp.SMPL(full_model, dtype=np.float64)
    b = {k: v.astype(np.float64) for k, v in inp.items()}
    # ---- the loss of the WHOLE batch: oracle forward over all 4096 bodies (chunked)
    joints = _oracle_joints(o, b["beta"], b["theta"])
    kp = onp.batch_orth_proj_idrot(joints, b["cam"])
    num, cnt = onp.kp_loss_parts(b["kp_gt"], kp)
    assert int(out["loss_parts"][1]) == cnt                                   # bit-exact count
    assert abs(out["loss_parts"][0] - num) < TOL * num
    assert abs(out["loss_parts"][3] - w_kp * num / cnt) < TOL * w_kp * num / cnt
    assert rel_err(out["joints"], joints) < TOL and rel_err(out["kp_pred"], kp) < TOL
    # ---- gradients of 64 sampled bodies (first, all-invisible, last, 61 random), scaled by the GLOBAL count
    idx = np.unique(np.concatenate([[0, 1, 2, B - 1], np.random.default_rng(3).choice(B, 64, replace=False)]))[:68]
    verts, j_s, Rs = o(b["beta"][idx], b["theta"][idx], get_skin=True)
    assert rel_err(out["verts"][idx], verts) < TOL and rel_err(out["Rs"][idx], Rs) < TOL
    kp_s = kp[idx]
    g3 = b["kp_gt"][idx]
    d_kp = w_kp * g3[:, :, 2:3] * np.sign(kp_s - g3[:, :, :2]) / cnt           # ops.py:35-47 differentiated, global count
    dj, dcam = onp.orth_proj_backward(j_s, b["cam"][idx], d_kp)
    db, dth = onp.smpl_backward(o, b["beta"][idx], b["theta"][idx], None, dj, None)
    assert rel_err(out["d_beta"][idx], db) < TOL
    assert rel_err(out["d_theta"][idx], dth) < TOL
    assert rel_err(out["d_cam"][idx], dcam) < TOL
    assert not out["d_theta"][1].any() and not out["d_cam"][1].any()          # all-invisible body: exactly zero


def _nn_is_valid(A, Bp, iab, iba, tol_d2):
    """Every index is an argmin of the exact squared distance up to tol_d2 (the rounding error of the
    reference's own fp32 expansion -2ab + |a|^2 + |b|^2 at pixel magnitudes)."""
    a2 = np.sum(A * A, 1)[:, None]
    b2 = np.sum(Bp * Bp, 1)[None, :]
    best_ba = np.full(Bp.shape[0], np.inf)
    worst_ab = 0.0
    for s in range(0, A.shape[0], 2048):
        d = -2.0 * A[s:s + 2048] @ Bp.T + a2[s:s + 2048] + b2
        rows = np.arange(d.shape[0])
        worst_ab = max(worst_ab, float(np.max(d[rows, iab[s:s + 2048]] - d.min(1))))
        best_ba = np.minimum(best_ba, d.min(0))
    d_ba = np.sum((Bp - A[iba]) ** 2, 1)
    return worst_ab <= tol_d2 and float(np.max(d_ba - best_ba)) <= tol_d2


def test_config3_like_mesh_step_vs_oracle(full_model):
    """16 images, P_i ~ 5-13 k pixels (SURVEY section 8d silhouettes), V = 6890, keypoint + mesh loss."""
    B, V = 16, 6890
    w_kp, w_mesh = 60.0, 0.001
    s = SMPL(full_model, max_batch=B)
    inp = synthetic.make_inputs(B, seed=31)
    seg = synthetic.make_silhouettes(B, seed=32)
    pts3 = synthetic.silhouette_points(seg)
    pts, offs = ops.silhouette_csr(pts3, B)
    assert pts.shape[0] > 8000 * (B - 1) and offs[3] == offs[2]               # ~10 k pixels per image, image 2 empty
    out = s.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], silhouette=(pts, offs), w_kp=w_kp, w_mesh=w_mesh)
    sil_gpu = projection.reproject_vertices(out["verts"], inp["cam"], [224.0, 224.0])
    _, _, iab, iba = ops._mesh_call(s.ctx, pts, offs, sil_gpu, False, True)

    o = onp.SMPL(full_model, dtype=np.float64)
    b = {k: v.astype(np.float64) for k, v in inp.items()}
    verts, joints, Rs = o(b["beta"], b["theta"], get_skin=True)
    assert rel_err(out["verts"], verts) < TOL
    sp = onp.reproject_vertices(verts, b["cam"], [224.0, 224.0])
    assert rel_err(sil_gpu, sp) < TOL
    denom = 3 + V
    mesh_idx = 0.0          # oracle value with the GPU's neighbour choice
    g_sp = np.zeros_like(sp)
    # The loss is |.|_1 + |.|_2 of differences to the nearest neighbour: its gradient is sign(a - b) and
    # (b - a) / |b - a|, discontinuous where a difference crosses zero.  With ~110 k vertices and ~160 k
    # pixels a few differences are below the fp32 resolution of a pixel coordinate (224 * 2^-23 * O(10)),
    # where the sign / direction is decided by rounding, in the reference's fp32 graph as much as here.
    # So the gradient of the loss w.r.t. silhouette_pred is evaluated in fp64 AT the GPU's silhouette_pred
    # (which matches the oracle's to 1e-4 above), and everything after it is the oracle's fp64 chain.
    spg = np.asarray(sil_gpu, dtype=np.float64)
    for i in range(B):
        A = pts[offs[i]:offs[i + 1]].astype(np.float64)
        if A.shape[0] == 0:
            assert (iba[i] == -1).all()
            continue
        ia, ib = iab[offs[i]:offs[i + 1]].astype(np.int64), iba[i].astype(np.int64)
        # (1) the indices are nearest neighbours within the rounding of the fp32 expansion: |p|^2 <= 1e5, ulp 2^-7
        assert _nn_is_valid(A, sp[i], ia, ib, tol_d2=0.06), "image %d: an index is not a nearest neighbour" % i
        # (2) bidirectional_dist and its gradient for those indices (argmin has no gradient, ops.py:68-69)
        diff = sp[i] - A[ib]
        mesh_idx += (np.sum(np.sqrt(np.sum(diff * diff, 1))) + np.sum(np.abs(A - sp[i][ia]))) / denom
        diff = spg[i] - A[ib]
        gB = diff / np.sqrt(np.sum(diff * diff, 1))[:, None]
        np.add.at(gB, ia, -np.sign(A - spg[i][ia]))
        g_sp[i] = gB / denom
    mesh_own = onp.mesh_reprojection_loss(pts3.astype(np.float64), sp, B)    # the oracle's own fp64 argmin
    assert abs(out["loss_parts"][2] - mesh_idx) < TOL * mesh_idx
    assert abs(out["loss_parts"][2] - mesh_own) < TOL * mesh_own
    kp = onp.batch_orth_proj_idrot(joints, b["cam"])
    num, cnt = onp.kp_loss_parts(b["kp_gt"], kp)
    total = w_kp * num / cnt + w_mesh * mesh_idx
    assert int(out["loss_parts"][1]) == cnt and abs(out["loss_parts"][3] - total) < TOL * total
    dj, dcam_kp = onp.orth_proj_backward(joints, b["cam"], w_kp * onp.kp_loss_backward(b["kp_gt"], kp))
    dv, dcam_mesh = onp.reproject_vertices_backward(verts, b["cam"], [224.0, 224.0], w_mesh * g_sp)
    db, dth = onp.smpl_backward(o, b["beta"], b["theta"], dv, dj, None)
    assert rel_err(out["d_beta"], db) < TOL
    assert rel_err(out["d_theta"], dth) < TOL
    assert rel_err(out["d_cam"], dcam_kp + dcam_mesh) < TOL


def test_config5_b65536_forward_spot_check(full_model):
    """The largest batch of the inference sweep: verts stay on the device (5.4 GB), 32 bodies are copied
    back and compared; joints and Rs of all bodies must be finite and the batch deterministic."""
    B, V = 65536, 6890
    s = SMPL(full_model, max_batch=B)
    inp = synthetic.make_inputs(B, seed=65)
    ctx = s.ctx
    d_beta, d_theta = ctx.to_device(inp["beta"]), ctx.to_device(inp["theta"])
    verts, joints, Rs = s(d_beta, d_theta, get_skin=True)
    idx = np.unique(np.concatenate([[0, 1, B - 1], np.random.default_rng(1).choice(B, 29, replace=False)]))
    got = np.empty((len(idx), V, 3), np.float32)
    for n, i in enumerate(idx):
        check(lib().smplb_memcpy_d2h(ctx.handle, got[n].ctypes.data, C.c_void_p(verts.ptr + int(i) * V * 12), V * 12))
    ctx.sync()
    j, R = joints.numpy(), Rs.numpy()
    assert np.isfinite(j).all() and np.isfinite(R).all()
    o = onp.SMPL(full_model, dtype=np.float64)
    v64, j64, R64 = o(inp["beta"][idx].astype(np.float64), inp["theta"][idx].astype(np.float64), get_skin=True)
    assert rel_err(got, v64) < TOL and rel_err(j[idx], j64) < TOL and rel_err(R[idx], R64) < TOL
    # per-sample results do not depend on the batch they are computed in: the same bodies at B = 32
    v2, j2, _ = s(inp["beta"][idx], inp["theta"][idx], get_skin=True)
    assert np.array_equal(v2, got) and np.array_equal(j2, j[idx])


def test_lsp_k14_on_the_6890_vertex_model(full_model):
    B = 64
    s = SMPL(full_model, joint_type="lsp", max_batch=B)
    inp = synthetic.make_inputs(B, seed=14, num_keypoints=14)
    out = s.step(inp["beta"], inp["theta"], inp["cam"], inp["kp_gt"], w_kp=1.0)
    assert out["joints"].shape == (B, 14, 3)
    o = onp.SMPL(full_model, "lsp", dtype=np.float64)
    b = {k: v.astype(np.float64) for k, v in inp.items()}
    verts, joints, Rs = o(b["beta"], b["theta"], get_skin=True)
    kp = onp.batch_orth_proj_idrot(joints, b["cam"])
    num, cnt = onp.kp_loss_parts(b["kp_gt"], kp)
    assert rel_err(out["verts"], verts) < TOL and rel_err(out["joints"], joints) < TOL
    assert int(out["loss_parts"][1]) == cnt and abs(out["loss_parts"][3] - num / cnt) < TOL * num / cnt
    dj, dcam = onp.orth_proj_backward(joints, b["cam"], onp.kp_loss_backward(b["kp_gt"], kp))
    db, dth = onp.smpl_backward(o, b["beta"], b["theta"], None, dj, None)
    assert rel_err(out["d_beta"], db) < TOL and rel_err(out["d_theta"], dth) < TOL and rel_err(out["d_cam"], dcam) < TOL
    # the cocoplus model's first 14 keypoints are the lsp ones (batch_smpl.py:80-81)
    s19 = SMPL(full_model, max_batch=B)
    j19 = s19(inp["beta"], inp["theta"])
    assert rel_err(j19[:, :14], out["joints"]) < 1e-6
