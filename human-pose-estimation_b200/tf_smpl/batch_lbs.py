"""Util functions for SMPL on libsmplb.so, mirroring reference
src/tf_smpl/batch_lbs.py: batch_skew (:15), batch_rodrigues (:42),
batch_lrotmin (:67), batch_global_rigid_transformation (:91)."""
import numpy as np

from .. import runtime
from .._lib import check, lib
from ..ops import _ctx_for


def batch_skew(vec, batch_size=None):
    """vec is N x 3 -> N x 3 x 3 skew-symmetric matrices."""
    ctx = _ctx_for(vec)
    a = runtime.Args(ctx)
    N = int(vec.shape[0]) if batch_size is None else int(batch_size)
    pv = a.inp(vec, (N, 3))
    out, po = a.out((N, 3, 3))
    check(lib().smplb_skew(ctx.handle, N, pv, po, a.mem))
    return out


def batch_rodrigues(theta, name=None):
    """theta is N x 3 -> N x 3 x 3 rotation matrices."""
    ctx = _ctx_for(theta)
    a = runtime.Args(ctx)
    N = int(theta.shape[0])
    pt = a.inp(theta, (N, 3))
    out, po = a.out((N, 3, 3))
    check(lib().smplb_rodrigues(ctx.handle, N, pt, po, a.mem))
    return out


def batch_lrotmin(theta, name=None):
    """theta N x 72 -> N x 207: rotations of the 23 non-root joints minus I."""
    ctx = _ctx_for(theta)
    a = runtime.Args(ctx)
    N = int(theta.shape[0])
    pt = a.inp(theta, (N, 72))
    out, po = a.out((N, 207))
    check(lib().smplb_lrotmin(ctx.handle, N, pt, po, a.mem))
    return out


def batch_global_rigid_transformation(Rs, Js, parent, rotate_base=False):
    """Rs N x 24 x 3 x 3, Js N x 24 x 3, parent [24] -> (new_J N x 24 x 3,
    A N x 24 x 4 x 4)."""
    if rotate_base:
        # the reference never takes this path (batch_smpl.py:135 passes the default)
        raise NotImplementedError("rotate_base=True is unused by the reference and not implemented")
    ctx = _ctx_for(Rs, parents=np.asarray(parent))
    a = runtime.Args(ctx)
    N = int(Rs.shape[0])
    pR = a.inp(Rs, (N, 24, 3, 3))
    pJ = a.inp(Js, (N, 24, 3))
    new_J, pn = a.out((N, 24, 3))
    A, pA = a.out((N, 24, 4, 4))
    check(lib().smplb_global_rigid(ctx.handle, N, pR, pJ, pn, pA, a.mem))
    return new_J, A
