"""CPU restatement of the hot path in torch (fp32 or fp64), gradients by autograd.
TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).

Why a second restatement next to smpl_numpy.py: this one is the CPU *baseline* of bench.py on the GPU
box.  The reference is TensorFlow graph code whose CPU cost is dominated by batched matmuls; the
reference's own files executed under oracle/tf_shim (torch-CPU) run ~4x faster than the numpy port on
the same cores because torch threads the batched matmuls -- but the reference checkout cannot travel to
the GPU box.  This file states the same formulation line by line on torch ops (including the
materialised [B,6890,24] weight tile and [B,6890,4,4] transforms of batch_smpl.py:139-147), so the
baseline timed there is as fast as the reference's own code is here.  Pinned to smpl_numpy.py (and
through it to the golden vectors) in tests/test_oracle.py.

Follows: src/tf_smpl/batch_lbs.py:15-64,91-152; src/tf_smpl/batch_smpl.py:88-160;
src/tf_smpl/projection.py:23-33; src/ops.py:35-47.
"""
import numpy as np
import torch


def batch_skew(vec):
    """batch_lbs.py:15-39."""
    N = vec.shape[0]
    z = torch.zeros(N, dtype=vec.dtype)
    return torch.stack([z, -vec[:, 2], vec[:, 1], vec[:, 2], z, -vec[:, 0], -vec[:, 1], vec[:, 0], z], dim=1).reshape(N, 3, 3)


def batch_rodrigues(theta):
    """batch_lbs.py:42-64."""
    angle = torch.norm(theta + 1e-8, dim=1, keepdim=True)
    r = (theta / angle).unsqueeze(-1)
    angle = angle.unsqueeze(-1)
    c, s = torch.cos(angle), torch.sin(angle)
    outer = torch.matmul(r, r.transpose(1, 2))
    eyes = torch.eye(3, dtype=theta.dtype).unsqueeze(0)
    return c * eyes + (1 - c) * outer + s * batch_skew(r[:, :, 0])


def batch_global_rigid_transformation(Rs, Js, parent):
    """batch_lbs.py:91-152, rotate_base=False."""
    N = Rs.shape[0]
    Js4 = Js.unsqueeze(-1)

    def make_A(R, t):
        R_homo = torch.nn.functional.pad(R, (0, 0, 0, 1))
        t_homo = torch.cat([t, torch.ones(N, 1, 1, dtype=R.dtype)], dim=1)
        return torch.cat([R_homo, t_homo], dim=2)

    results = [make_A(Rs[:, 0], Js4[:, 0])]
    for i in range(1, parent.shape[0]):
        j_here = Js4[:, i] - Js4[:, parent[i]]
        results.append(torch.matmul(results[parent[i]], make_A(Rs[:, i], j_here)))
    results = torch.stack(results, dim=1)
    new_J = results[:, :, :3, 3]
    Js_w0 = torch.cat([Js4, torch.zeros(N, parent.shape[0], 1, 1, dtype=Rs.dtype)], dim=2)
    init_bone = torch.matmul(results, Js_w0)
    init_bone = torch.nn.functional.pad(init_bone, (3, 0))
    return new_J, results - init_bone


class SMPL(object):
    """batch_smpl.py:25-160 on a model dict with the reference pickle's keys."""

    def __init__(self, model, joint_type="cocoplus", dtype=torch.float32):
        und = lambda x: np.asarray(x if isinstance(x, np.ndarray) else x.r)   # noqa: E731
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dtype)   # noqa: E731
        self.dtype = dtype
        self.v_template = t(und(model["v_template"]))
        self.size = [self.v_template.shape[0], 3]
        self.num_betas = model["shapedirs"].shape[-1]
        self.shapedirs = t(np.reshape(und(model["shapedirs"]), [-1, self.num_betas]).T)
        self.J_regressor = t(np.asarray(model["J_regressor"].T.todense()))
        npb = model["posedirs"].shape[-1]
        self.posedirs = t(np.reshape(und(model["posedirs"]), [-1, npb]).T)
        self.parents = model["kintree_table"][0].astype(np.int32)
        self.weights = t(und(model["weights"]))
        jr = np.asarray(model["cocoplus_regressor"].T.todense())
        self.joint_regressor = t(jr[:, :14] if joint_type == "lsp" else jr)

    def __call__(self, beta, theta, get_skin=False):
        B, V = beta.shape[0], self.size[0]
        v_shaped = torch.matmul(beta, self.shapedirs).reshape(-1, V, 3) + self.v_template            # :110-112
        J = torch.stack([torch.matmul(v_shaped[:, :, c], self.J_regressor) for c in range(3)], dim=2)  # :115-118
        Rs = batch_rodrigues(theta.reshape(-1, 3)).reshape(-1, 24, 3, 3)                             # :122-123
        pose_feature = (Rs[:, 1:] - torch.eye(3, dtype=self.dtype)).reshape(-1, 207)                 # :126-127
        v_posed = torch.matmul(pose_feature, self.posedirs).reshape(-1, V, 3) + v_shaped             # :130-132
        self.J_transformed, A = batch_global_rigid_transformation(Rs, J, self.parents)               # :135
        W = self.weights.repeat(B, 1).reshape(B, -1, 24)                                             # :139-140 (tf.tile)
        T = torch.matmul(W, A.reshape(B, 24, 16)).reshape(B, -1, 4, 4)                               # :142-144
        v_posed_homo = torch.cat([v_posed, torch.ones(B, V, 1, dtype=self.dtype)], dim=2)            # :145-146
        v_homo = torch.matmul(T, v_posed_homo.unsqueeze(-1))                                         # :147
        verts = v_homo[:, :, :3, 0]                                                                  # :149
        joints = torch.stack([torch.matmul(verts[:, :, c], self.joint_regressor) for c in range(3)], dim=2)  # :152-155
        if get_skin:
            return verts, joints, Rs
        return joints


def batch_orth_proj_idrot(X, camera):
    """projection.py:23-33."""
    camera = camera.reshape(-1, 1, 3)
    X_trans = X[:, :, :2] + camera[:, :, 1:]
    shape = X_trans.shape
    return (camera[:, :, 0] * X_trans.reshape(shape[0], -1)).reshape(shape)


def kp_reprojection_loss(kp_gt, kp_pred):
    """ops.py:35-47: absolute_difference with weights vis, reduction SUM_BY_NONZERO_WEIGHTS."""
    kp_gt = kp_gt.reshape(-1, 3)
    kp_pred = kp_pred.reshape(-1, 2)
    vis = kp_gt[:, 2:3]
    num = torch.sum(torch.abs(kp_pred - kp_gt[:, :2]) * vis)
    cnt = 2 * int(torch.count_nonzero(vis))
    return num / cnt if cnt > 0 else num * 0


def step(smpl, beta, theta, cam, kp_gt):
    """BASELINE config 2 on numpy inputs: forward, keypoint loss, autograd backward.
    Returns (verts, joints, Rs, kp_pred, loss, d_beta, d_theta, d_cam) as numpy."""
    t = lambda a, g: torch.from_numpy(np.ascontiguousarray(a)).to(smpl.dtype).requires_grad_(g)   # noqa: E731
    b, th, c, g = t(beta, True), t(theta, True), t(cam, True), t(kp_gt, False)
    verts, joints, Rs = smpl(b, th, get_skin=True)
    kp = batch_orth_proj_idrot(joints, c)
    loss = kp_reprojection_loss(g, kp)
    if loss.requires_grad and float(loss.detach()) != 0.0:
        db, dth, dc = torch.autograd.grad(loss, [b, th, c])
    else:
        db, dth, dc = torch.zeros_like(b), torch.zeros_like(th), torch.zeros_like(c)
    n = lambda x: x.detach().numpy()   # noqa: E731
    return n(verts), n(joints), n(Rs), n(kp), float(loss), n(db), n(dth), n(dc)
