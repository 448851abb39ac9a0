"""Oracle B -- numpy restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

This file is the CPU checker for the CUDA path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs
may import it; the product package never does (and has no CPU fallback).

Every function follows the cited lines of the reference (paths relative to
/root/reference).  The arithmetic itself lives in TensorFlow (pinned
tensorflow==2.0.0-beta1, README.md:67,71 -- un-vendored and not installable
here); the published semantics of the tf ops used are restated.

Pinning: the reference ships NO tests, golden vectors or fixtures for this path
(SURVEY.md §4), so the pin is the reference's own code executed here:
oracle/run_reference.py runs the four reference files unchanged under
oracle/tf_shim and tests/test_oracle.py checks this restatement against it
(and against the committed outputs in tests/golden/ made by
oracle/make_golden.py from that run).

`dtype` selects float64 (checker for tolerances) or float32 (CPU baseline that
mirrors the reference's fp32 formulation, including its materialised
[B,6890,24] weight tile and [B,6890,4,4] transforms).
"""
import numpy as np


# --------------------------------------------------------------------------
# src/tf_smpl/batch_lbs.py
# --------------------------------------------------------------------------
def batch_skew(vec):
    """batch_lbs.py:15-39 -- [[0,-z,y],[z,0,-x],[-y,x,0]] (scatter at flat
    indices 1,2,3,5,6,7)."""
    N = vec.shape[0]
    res = np.zeros((N, 9), dtype=vec.dtype)
    res[:, 1] = -vec[:, 2]
    res[:, 2] = vec[:, 1]
    res[:, 3] = vec[:, 2]
    res[:, 5] = -vec[:, 0]
    res[:, 6] = -vec[:, 1]
    res[:, 7] = vec[:, 0]
    return res.reshape(N, 3, 3)


def batch_rodrigues(theta):
    """batch_lbs.py:42-64 -- angle = ||theta + 1e-8||, r = theta / angle,
    R = cos*I + (1-cos)*r r^T + sin*skew(r)."""
    dt = theta.dtype
    angle = np.sqrt(np.sum(np.square(theta + dt.type(1e-8)), axis=1))[:, None]
    r = (theta / angle)[:, :, None]
    angle = angle[:, :, None]
    c = np.cos(angle)
    s = np.sin(angle)
    outer = np.matmul(r, r.transpose(0, 2, 1))
    eyes = np.eye(3, dtype=dt)[None]
    return c * eyes + (1 - c) * outer + s * batch_skew(r[:, :, 0])


def batch_lrotmin(theta):
    """batch_lbs.py:67-88 (unused by the reference; API parity)."""
    theta = theta[:, 3:]
    Rs = batch_rodrigues(theta.reshape(-1, 3))
    return (Rs - np.eye(3, dtype=theta.dtype)).reshape(-1, 207)


def batch_global_rigid_transformation(Rs, Js, parent):
    """batch_lbs.py:91-152 with rotate_base=False: G_0=[R_0|J_0],
    G_i = G_parent(i) . [R_i | J_i - J_parent(i)]; new_J = G[:,:,:3,3];
    A = G - pad(G . [J;0])."""
    N = Rs.shape[0]
    dt = Rs.dtype
    Js4 = Js[:, :, :, None]

    def make_A(R, t):
        out = np.zeros((N, 4, 4), dtype=dt)
        out[:, :3, :3] = R
        out[:, :3, 3:4] = t
        out[:, 3, 3] = 1
        return out

    results = [make_A(Rs[:, 0], Js4[:, 0])]
    for i in range(1, parent.shape[0]):
        j_here = Js4[:, i] - Js4[:, parent[i]]
        results.append(np.matmul(results[parent[i]], make_A(Rs[:, i], j_here)))
    results = np.stack(results, axis=1)
    new_J = results[:, :, :3, 3]
    Js_w0 = np.concatenate([Js4, np.zeros((N, parent.shape[0], 1, 1), dtype=dt)], axis=2)
    init_bone = np.matmul(results, Js_w0)
    init_bone = np.pad(init_bone, [[0, 0], [0, 0], [0, 0], [3, 0]])
    return new_J, results - init_bone


# --------------------------------------------------------------------------
# src/tf_smpl/batch_smpl.py
# --------------------------------------------------------------------------
class SMPL(object):
    """batch_smpl.py:25-160 on a model dict with the reference pickle's keys."""

    def __init__(self, model, joint_type="cocoplus", dtype=np.float64):
        dt = np.dtype(dtype)
        self.dtype = dt
        und = lambda x: np.asarray(x if isinstance(x, np.ndarray) else x.r)
        self.v_template = und(model["v_template"]).astype(dt)                       # :34-38
        self.size = [self.v_template.shape[0], 3]
        self.num_betas = model["shapedirs"].shape[-1]
        self.shapedirs = np.reshape(und(model["shapedirs"]), [-1, self.num_betas]).T.astype(dt)   # :44-47
        self.J_regressor = np.asarray(model["J_regressor"].T.todense()).astype(dt)  # :50-54
        npb = model["posedirs"].shape[-1]
        self.posedirs = np.reshape(und(model["posedirs"]), [-1, npb]).T.astype(dt)  # :57-62
        self.parents = model["kintree_table"][0].astype(np.int32)                   # :65
        self.weights = und(model["weights"]).astype(dt)                             # :68-72
        self.joint_regressor = np.asarray(model["cocoplus_regressor"].T.todense()).astype(dt)  # :75-79
        if joint_type == "lsp":
            self.joint_regressor = self.joint_regressor[:, :14]                     # :80-81
        if joint_type not in ("cocoplus", "lsp"):
            raise ValueError("joint_type must be cocoplus or lsp")

    def __call__(self, beta, theta, get_skin=False, intermediates=None):
        dt = self.dtype
        beta = np.asarray(beta, dtype=dt)
        theta = np.asarray(theta, dtype=dt)
        B = beta.shape[0]
        V = self.size[0]
        v_shaped = np.matmul(beta, self.shapedirs).reshape(-1, V, 3) + self.v_template       # :110-112
        J = np.stack([np.matmul(v_shaped[:, :, c], self.J_regressor) for c in range(3)], axis=2)  # :115-118
        Rs = batch_rodrigues(theta.reshape(-1, 3)).reshape(-1, 24, 3, 3)                     # :122-123
        pose_feature = (Rs[:, 1:] - np.eye(3, dtype=dt)).reshape(-1, 207)                    # :126-127
        v_posed = np.matmul(pose_feature, self.posedirs).reshape(-1, V, 3) + v_shaped        # :130-132
        self.J_transformed, A = batch_global_rigid_transformation(Rs, J, self.parents)       # :135
        W = np.tile(self.weights, (B, 1)).reshape(B, -1, 24)                                 # :139-140
        T = np.matmul(W, A.reshape(B, 24, 16)).reshape(B, -1, 4, 4)                          # :142-144
        v_posed_homo = np.concatenate([v_posed, np.ones((B, V, 1), dtype=dt)], axis=2)       # :145-146
        v_homo = np.matmul(T, v_posed_homo[:, :, :, None])                                   # :147
        verts = v_homo[:, :, :3, 0]                                                          # :149
        joints = np.stack([np.matmul(verts[:, :, c], self.joint_regressor) for c in range(3)], axis=2)  # :152-155
        if intermediates is not None:
            intermediates.update(v_shaped=v_shaped, J=J, pose_feature=pose_feature, v_posed=v_posed, A=A)
        if get_skin:
            return verts, joints, Rs
        return joints


# --------------------------------------------------------------------------
# src/tf_smpl/projection.py
# --------------------------------------------------------------------------
def batch_orth_proj_idrot(X, camera):
    """projection.py:23-33 -- s * (X[:, :, :2] + t)."""
    camera = camera.reshape(-1, 1, 3)
    X_trans = X[:, :, :2] + camera[:, :, 1:]
    shp = X_trans.shape
    return (camera[:, :, 0] * X_trans.reshape(shp[0], -1)).reshape(shp)


def reproject_vertices(verts, cam, im_size):
    """projection.py:45-56 -- ((proj + 1) * 0.5) * im_size."""
    p = batch_orth_proj_idrot(verts, cam)
    return ((p + np.ones_like(p)) * p.dtype.type(0.5)) * np.asarray(im_size, dtype=p.dtype)


# --------------------------------------------------------------------------
# src/ops.py
# --------------------------------------------------------------------------
def kp_loss_parts(kp_gt, kp_pred):
    """Numerator and integer count of ops.py:35-47 (tf.compat.v1.losses.
    absolute_difference, reduction SUM_BY_NONZERO_WEIGHTS, weights [BK,1]
    broadcast to the [BK,2] loss)."""
    kp_gt = kp_gt.reshape(-1, 3)
    kp_pred = kp_pred.reshape(-1, 2)
    vis = kp_gt[:, 2:3].astype(kp_pred.dtype)
    num = np.sum(np.abs(kp_pred - kp_gt[:, :2]) * vis)
    cnt = int(np.count_nonzero(vis)) * 2
    return num, cnt


def kp_reprojection_loss(kp_gt, kp_pred, scale=1.0):
    num, cnt = kp_loss_parts(kp_gt, kp_pred)
    return num / cnt if cnt > 0 else num * 0


def find_nearest_neighbors(A, B, chunk=2048):
    """ops.py:60-71 -- d2 = -2 A B^T + |A|^2 + |B|^2, argmin along both axes
    (first index on ties).  Chunked over A so fp64 runs fit in memory."""
    nA, nB = A.shape[0], B.shape[0]
    a2 = np.sum(np.square(A), 1)[:, None]
    b2 = np.sum(np.square(B), axis=1)[None, :]
    ind_AB = np.empty(nA, dtype=np.int64)
    best_BA = np.full(nB, np.inf, dtype=A.dtype)
    ind_BA = np.zeros(nB, dtype=np.int64)
    for s in range(0, nA, chunk):
        d = A.dtype.type(-2.0) * np.matmul(A[s:s + chunk], B.T) + a2[s:s + chunk] + b2
        ind_AB[s:s + chunk] = np.argmin(d, 1)
        loc = np.argmin(d, 0)
        val = d[loc, np.arange(nB)]
        upd = val < best_BA                       # strict: keep the first index on ties
        best_BA[upd] = val[upd]
        ind_BA[upd] = loc[upd] + s
    return ind_AB, ind_BA


def bidirectional_dist(A, B):
    """ops.py:83-102 -- sum_b ||B_b - A[nn]||_2 + sum_a ||A_a - B[nn]||_1."""
    ind_AB, ind_BA = find_nearest_neighbors(A, B)
    dist_BA = np.sqrt(np.sum(np.square(B - A[ind_BA]), axis=1))
    dist_AB = np.sum(np.abs(A - B[ind_AB]), axis=1)
    return np.sum(dist_BA) + np.sum(dist_AB)


def mesh_reprojection_loss(silhouette_gt, silhouette_pred, batch_size):
    """ops.py:117-137 -- per image: rows with col0 == i, point = (col2, col1);
    bidirectional_dist / (3 + 6890); summed over the batch.  An image with no
    silhouette pixels is undefined in TF (argmin over an empty axis); it
    contributes 0 here."""
    dt = silhouette_pred.dtype
    denom = silhouette_gt.shape[1] + silhouette_pred.shape[1]
    loss = dt.type(0)
    for i in range(batch_size):
        rows = silhouette_gt[silhouette_gt[:, 0] == i]
        if rows.shape[0] == 0:
            continue
        pts = np.stack([rows[:, 2], rows[:, 1]], axis=1).astype(dt)
        loss = loss + bidirectional_dist(pts, silhouette_pred[i]) / denom
    return loss


def compute_gradient_penalty(gradients):
    """ops.py:153-172 -- sum_i (1 - ||mean_axis0 g_i||_F)^2."""
    p = 0
    for g in gradients:
        p = p + np.square(1.0 - np.sqrt(np.sum(np.square(np.mean(g, axis=0)))))
    return p


# --------------------------------------------------------------------------
# Backward restatements (the reference relies on TF autodiff,
# src/trainer.py:383,502; formulas in SURVEY.md appendix B, validated against
# torch autograd through the shim-run reference in tests/test_oracle.py).
# --------------------------------------------------------------------------
def rodrigues_backward(theta, dR):
    """theta [N,3], dR [N,3,3] -> d_theta [N,3]."""
    dt = theta.dtype
    te = theta + dt.type(1e-8)
    a = np.sqrt(np.sum(te * te, axis=1))
    r = theta / a[:, None]
    c, s = np.cos(a), np.sin(a)
    G = dR
    ax = np.stack([G[:, 2, 1] - G[:, 1, 2], G[:, 0, 2] - G[:, 2, 0], G[:, 1, 0] - G[:, 0, 1]], axis=1)
    Gr = np.einsum("nij,nj->ni", G, r)
    GTr = np.einsum("nji,nj->ni", G, r)
    g_c = np.trace(G, axis1=1, axis2=2) - np.sum(r * Gr, axis=1)
    g_s = np.sum(r * ax, axis=1)
    g_r = (1 - c)[:, None] * (Gr + GTr) + s[:, None] * ax
    g_a = -s * g_c + c * g_s
    u = te / a[:, None]
    return g_r / a[:, None] - (np.sum(theta * g_r, axis=1) / (a * a))[:, None] * u + g_a[:, None] * u


def smpl_backward(smpl, beta, theta, d_verts=None, d_joints=None, d_Rs=None):
    """Gradients of sum(verts*d_verts) + sum(joints*d_joints) + sum(Rs*d_Rs)
    w.r.t. beta [B,10] and theta [B,72]."""
    dt = smpl.dtype
    beta = np.asarray(beta, dtype=dt)
    theta = np.asarray(theta, dtype=dt).reshape(beta.shape[0], 72)
    B = beta.shape[0]
    V = smpl.size[0]
    inter = {}
    verts, joints, Rs = smpl(beta, theta, get_skin=True, intermediates=inter)
    J, v_posed, A = inter["J"], inter["v_posed"], inter["A"]
    parents = smpl.parents
    g = np.zeros((B, V, 3), dtype=dt)
    if d_verts is not None:
        g = g + np.asarray(d_verts, dtype=dt)
    if d_joints is not None:
        g = g + np.matmul(smpl.joint_regressor, np.asarray(d_joints, dtype=dt))       # [V,K] @ [B,K,3]
    W = smpl.weights
    AR, At = A[:, :, :3, :3], A[:, :, :3, 3]
    # T = W A, verts = T [p;1]  =>  dT = g (x) [p;1],  dA = W^T dT,  dp = T_R^T g
    ph = np.concatenate([v_posed, np.ones((B, V, 1), dtype=dt)], axis=2)
    dT = (g[:, :, :, None] * ph[:, :, None, :]).reshape(B, V, 12)
    dA = np.matmul(W.T, dT).reshape(B, 24, 3, 4)
    dAR, dAt = dA[:, :, :, :3], dA[:, :, :, 3]
    TR = np.matmul(W, np.ascontiguousarray(AR).reshape(B, 24, 9)).reshape(B, V, 3, 3)
    dp = np.sum(TR * g[:, :, :, None], axis=2)
    # global transforms Rg, tg from A: A_R = Rg, A_t = tg - Rg J
    Rg = AR
    dRg = dAR - np.einsum("bjr,bjc->bjrc", dAt, J)
    dtg = dAt.copy()
    dJ = -np.einsum("bjrc,bjr->bjc", Rg, dAt)
    dR = np.zeros((B, 24, 3, 3), dtype=dt)
    for i in range(23, 0, -1):
        p = parents[i]
        jrel = J[:, i] - J[:, p]
        dR[:, i] = np.einsum("bkr,bkc->brc", Rg[:, p], dRg[:, i])
        dRg[:, p] += np.einsum("brk,bck->brc", dRg[:, i], Rs[:, i]) + np.einsum("br,bc->brc", dtg[:, i], jrel)
        djr = np.einsum("bkr,bk->br", Rg[:, p], dtg[:, i])
        dJ[:, i] += djr
        dJ[:, p] -= djr
        dtg[:, p] += dtg[:, i]
    dR[:, 0] = dRg[:, 0]
    dJ[:, 0] += dtg[:, 0]
    # blend shapes
    dp_flat = dp.reshape(B, V * 3)
    d_pf = np.matmul(dp_flat, smpl.posedirs.T)
    dR[:, 1:] += d_pf.reshape(B, 23, 3, 3)
    if d_Rs is not None:
        dR = dR + np.asarray(d_Rs, dtype=dt).reshape(B, 24, 3, 3)
    # v_shaped receives dp (through v_posed) and J_regressor^T dJ
    dvs = dp + np.matmul(smpl.J_regressor, dJ)                                        # [V,24] @ [B,24,3]
    d_beta = np.matmul(dvs.reshape(B, V * 3), smpl.shapedirs.T)
    d_theta = rodrigues_backward(theta.reshape(-1, 3), dR.reshape(-1, 3, 3)).reshape(B, 72)
    return d_beta, d_theta


def orth_proj_backward(X, camera, d_out):
    """d_out [B,N,2] -> dX [B,N,3] (z gets 0), d_cam [B,3]."""
    s = camera[:, 0][:, None, None]
    t = camera[:, 1:][:, None, :]
    dX = np.zeros_like(X)
    dX[:, :, :2] = s * d_out
    ds = np.sum(d_out * (X[:, :, :2] + t), axis=(1, 2))
    dtr = camera[:, 0][:, None] * np.sum(d_out, axis=1)
    return dX, np.concatenate([ds[:, None], dtr], axis=1)


def kp_loss_backward(kp_gt, kp_pred):
    """d loss / d kp_pred: vis * sign(pred - gt) / num_present."""
    shp = kp_pred.shape
    g3 = kp_gt.reshape(-1, 3)
    p2 = kp_pred.reshape(-1, 2)
    vis = g3[:, 2:3].astype(p2.dtype)
    cnt = int(np.count_nonzero(vis)) * 2
    if cnt == 0:
        return np.zeros(shp, dtype=p2.dtype)
    return (vis * np.sign(p2 - g3[:, :2]) / cnt).reshape(shp)


def mesh_loss_backward(silhouette_gt, silhouette_pred, batch_size):
    """d loss / d silhouette_pred [B,V,2] (indices treated as constants, as TF's
    argmin has no gradient)."""
    dt = silhouette_pred.dtype
    denom = silhouette_gt.shape[1] + silhouette_pred.shape[1]
    out = np.zeros_like(silhouette_pred)
    for i in range(batch_size):
        rows = silhouette_gt[silhouette_gt[:, 0] == i]
        if rows.shape[0] == 0:
            continue
        A = np.stack([rows[:, 2], rows[:, 1]], axis=1).astype(dt)
        Bp = silhouette_pred[i]
        ind_AB, ind_BA = find_nearest_neighbors(A, Bp)
        diff = Bp - A[ind_BA]
        nrm = np.sqrt(np.sum(diff * diff, axis=1))[:, None]
        gB = diff / nrm
        np.add.at(gB, ind_AB, -np.sign(A - Bp[ind_AB]))
        out[i] = gB / denom
    return out


def reproject_vertices_backward(verts, cam, im_size, d_out):
    im = np.asarray(im_size, dtype=verts.dtype)
    return orth_proj_backward(verts, cam, d_out * (verts.dtype.type(0.5) * im))


def gradient_penalty_backward(gradients):
    """d penalty / d g_i = -2 (1-n_i) * mean_i / (n_i * M)."""
    out = []
    for g in gradients:
        m = np.mean(g, axis=0)
        n = np.sqrt(np.sum(m * m))
        out.append(np.broadcast_to(-2.0 * (1.0 - n) * m / (n * g.shape[0]), g.shape).astype(g.dtype))
    return out


# --------------------------------------------------------------------------
# src/models.py (SURVEY.md section 8f neighbours of the path)
# --------------------------------------------------------------------------
def precompute_C_matrix(num_joints=14):
    """models.py:97-118."""
    num_bones = num_joints - 1
    C = np.zeros([num_joints, num_bones])
    C[np.arange(num_bones), np.arange(num_bones)] = 1
    C[np.array([1, 2, 8, 9, 3, 4, 7, 8, 12, 12, 9, 10, 13]), np.arange(num_bones)] = -1
    return C


def get_kcs(joints, C_matrix, num_joints=14):
    """models.py:123-139 -- B = joints[:, :14]^T C per sample, KCS = B^T B (the reference
    forms an N x 13 x 13 x N tensor and takes its diagonal; same numbers)."""
    j = joints[:, :num_joints, :]
    Bm = np.einsum("njc,ja->nca", j, C_matrix)
    return np.einsum("nca,ncb->nab", Bm, Bm)


def get_kcs_backward(joints, C_matrix, d_kcs, num_joints=14):
    j = joints[:, :num_joints, :]
    Bm = np.einsum("njc,ja->nca", j, C_matrix)
    dB = np.einsum("nab,ncb->nca", d_kcs + d_kcs.transpose(0, 2, 1), Bm)
    out = np.zeros_like(joints)
    out[:, :num_joints, :] = np.einsum("ja,nca->njc", C_matrix, dB)
    return out
