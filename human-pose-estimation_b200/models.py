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
    """fake + alpha * (real - fake) with one alpha per sample (trainer.py:551-557)."""
    ctx = _ctx_for(fake)
    a = runtime.Args(ctx)
    N = int(fake.shape[0])
    row = int(np.prod(fake.shape[1:]))
    pf, pr, pa = a.inp(fake, (N, row)), a.inp(real, (N, row)), a.inp(alpha, (N,))
    out, po = a.out(tuple(fake.shape))
    check(lib().smplb_interpolate(ctx.handle, N, row, pf, pr, pa, po, a.mem))
    return out
