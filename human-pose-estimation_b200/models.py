"""The pieces of reference src/models.py that sit next to the hot path:
precompute_C_matrix (:97-118) and get_kcs (:123-139) on libsmplb.so, plus the critic-input
interpolation of src/trainer.py:551-557.  The networks themselves are out of scope."""
import numpy as np

from . import runtime
from ._lib import check, lib
from .ops import _ctx_for


def precompute_C_matrix(num_joints=14):
    """Bone matrix C [14,13]: column b has +1 at joint b and -1 at the joint the bone ends in
    (RepNet's kinematic chain space)."""
    assert num_joints == 14, "num_joints must be 14 for now."
    num_bones = num_joints - 1
    ends = np.array([1, 2, 8, 9, 3, 4, 7, 8, 12, 12, 9, 10, 13])
    C = np.zeros([num_joints, num_bones], dtype=np.float32)
    C[np.arange(num_bones), np.arange(num_bones)] = 1
    C[ends, np.arange(num_bones)] = -1
    return C


def get_kcs(joints, C_matrix, num_joints=14):
    """joints N x K x 3 (first 14 used), C 14 x 13 -> KCS N x 13 x 13 (B^T B per sample)."""
    assert num_joints == 14
    ctx = _ctx_for(joints)
    a = runtime.Args(ctx)
    N, K = int(joints.shape[0]), int(joints.shape[1])
    pj = a.inp(joints, (N, K, 3))
    C = C_matrix.numpy() if isinstance(C_matrix, runtime.DeviceArray) else np.asarray(C_matrix, dtype=np.float32)
    pc = ctx.to_device(C).ptr if a.mem == runtime.DEVICE else a.inp(C, (14, 13))
    out, po = a.out((N, 13, 13))
    check(lib().smplb_kcs(ctx.handle, N, K, pj, pc, po, a.mem))
    return out


def get_kcs_backward(joints, C_matrix, d_kcs):
    """d_kcs N x 13 x 13 -> d_joints N x K x 3 (zero for the face keypoints)."""
    ctx = _ctx_for(joints)
    a = runtime.Args(ctx)
    N, K = int(joints.shape[0]), int(joints.shape[1])
    pj = a.inp(joints, (N, K, 3))
    pk = a.inp(d_kcs, (N, 13, 13))
    C = C_matrix.numpy() if isinstance(C_matrix, runtime.DeviceArray) else np.asarray(C_matrix, dtype=np.float32)
    pc = ctx.to_device(C).ptr if a.mem == runtime.DEVICE else a.inp(C, (14, 13))
    out, po = a.out((N, K, 3))
    check(lib().smplb_kcs_backward(ctx.handle, N, K, pj, pc, pk, po, a.mem))
    return out


def interpolate(fake, real, alpha):
    """fake + alpha * (real - fake) (trainer.py:551-557).  alpha has the tensor's own shape (the
    reference draws tf.random.uniform(x.shape), trainer.py:548-550) or one value per sample."""
    ctx = _ctx_for(fake)
    a = runtime.Args(ctx)
    N = int(fake.shape[0])
    row = int(np.prod(fake.shape[1:]))
    if int(np.prod(alpha.shape)) == N * row and row > 1:      # element-wise: rows of length 1
        N, row = N * row, 1
    pf, pr, pa = a.inp(fake, (N, row)), a.inp(real, (N, row)), a.inp(alpha, (N,))
    out, po = a.out(tuple(fake.shape))
    check(lib().smplb_interpolate(ctx.handle, N, row, pf, pr, pa, po, a.mem))
    return out


def critic_inputs(fake_joints, real_joints, alpha_joints, fake_shapes, real_shapes, alpha_shapes, fake_Rs, real_Rs,
                  alpha_Rs, C_matrix):
    """The interpolated inputs of the critic's gradient-penalty pass (trainer.py:548-557) in one launch:
    returns (joints_hat [N,K,3], kcs_hat [N,13,13], shapes_hat [N,10], Rs_hat [N,23,3,3]); alphas are element-wise."""
    ctx = _ctx_for(fake_joints)
    a = runtime.Args(ctx)
    N, K = int(fake_joints.shape[0]), int(fake_joints.shape[1])
    C = C_matrix.numpy() if isinstance(C_matrix, runtime.DeviceArray) else np.asarray(C_matrix, dtype=np.float32)
    p = [a.inp(x, (N, K, 3)) for x in (fake_joints, real_joints, alpha_joints)]
    p += [a.inp(x, (N, 10)) for x in (fake_shapes, real_shapes, alpha_shapes)]
    p += [a.inp(x, (N, 207)) for x in (fake_Rs, real_Rs, alpha_Rs)]
    pc = ctx.to_device(C).ptr if a.mem == runtime.DEVICE else a.inp(C, (14, 13))
    oj, pj = a.out((N, K, 3))
    ok, pk = a.out((N, 13, 13))
    os_, ps = a.out((N, 10))
    oR, pR = a.out((N, 23, 3, 3))
    check(lib().smplb_critic_inputs(ctx.handle, N, K, *p, pc, pj, pk, ps, pR, a.mem))
    return oj, ok, os_, oR


def critic_gradient_penalty(joints_hat, C_matrix, g_kcs, g_joints, g_shapes, g_Rs, m_total=None, want=()):
    """compute_gradient_penalty (ops.py:153-172) fused with the backward of get_kcs (trainer.py:566-572): the
    critic's partial derivatives in, the penalty out; `want` may name "col_sums" and / or "g_joints_total"."""
    ctx = _ctx_for(joints_hat)
    a = runtime.Args(ctx)
    M, K = int(joints_hat.shape[0]), int(joints_hat.shape[1])
    C = C_matrix.numpy() if isinstance(C_matrix, runtime.DeviceArray) else np.asarray(C_matrix, dtype=np.float32)
    pj = a.inp(joints_hat, (M, K, 3))
    pc = ctx.to_device(C).ptr if a.mem == runtime.DEVICE else a.inp(C, (14, 13))
    pg = [a.inp(g_kcs, (M, 169)), a.inp(g_joints, (M, 42)), a.inp(g_shapes, (M, 10)), a.inp(g_Rs, (M, 207))]
    pen, pp = a.out((1,))
    sums, psum = a.out((428,), want="col_sums" in want)
    gt, pgt = a.out((M, 14, 3), want="g_joints_total" in want)
    check(lib().smplb_critic_gradient_penalty(ctx.handle, M, K, int(m_total or M), pj, pc, *pg, pp, psum, pgt, a.mem))
    pv = pen.numpy()[0] if isinstance(pen, runtime.DeviceArray) else pen[0]
    res = {"penalty": np.float32(pv)}
    if sums is not None:
        res["col_sums"] = sums
    if gt is not None:
        res["g_joints_total"] = gt
    return res
