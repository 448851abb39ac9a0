"""Reprojection utilities on libsmplb.so, mirroring reference
src/tf_smpl/projection.py: batch_orth_proj_idrot (:23), reproject_vertices (:45)."""
import numpy as np

from .. import runtime
from .._lib import check, lib
from ..ops import _ctx_for


def batch_orth_proj_idrot(X, camera, name=None):
    """X: N x num_points x 3, camera: N x 3 -> N x num_points x 2."""
    ctx = _ctx_for(X)
    a = runtime.Args(ctx)
    N, n = int(X.shape[0]), int(X.shape[1])
    pX, pc = a.inp(X, (N, n, 3)), a.inp(camera, (N, 3))
    out, po = a.out((N, n, 2))
    check(lib().smplb_orth_proj(ctx.handle, N, n, pX, pc, po, a.mem))
    return out


def reproject_vertices(verts, cam, im_size, name=None):
    """verts: N x 6890 x 3, cam: N x 3, im_size: [w, h] -> pixel coords N x 6890 x 2."""
    ctx = _ctx_for(verts)
    a = runtime.Args(ctx)
    N, n = int(verts.shape[0]), int(verts.shape[1])
    im = np.asarray(im_size if not isinstance(im_size, runtime.DeviceArray) else im_size.numpy(), dtype=np.float32)
    pX, pc = a.inp(verts, (N, n, 3)), a.inp(cam, (N, 3))
    out, po = a.out((N, n, 2))
    check(lib().smplb_reproject_vertices(ctx.handle, N, n, pX, pc, float(im[0]), float(im[1]), po, a.mem))
    return out


def projection_backward(X, camera, d_out, im_size=None):
    """Backward of batch_orth_proj_idrot (im_size None) or reproject_vertices:
    d_out N x n x 2 -> (d_X N x n x 3, d_cam N x 3)."""
    ctx = _ctx_for(X)
    a = runtime.Args(ctx)
    N, n = int(X.shape[0]), int(X.shape[1])
    pX, pc, pd = a.inp(X, (N, n, 3)), a.inp(camera, (N, 3)), a.inp(d_out, (N, n, 2))
    dX, pdX = a.out((N, n, 3))
    dc, pdc = a.out((N, 3))
    pixel = 0 if im_size is None else 1
    im = np.asarray([0, 0] if im_size is None else im_size, dtype=np.float32)
    check(lib().smplb_proj_backward(ctx.handle, N, n, pX, pc, pd, pixel, float(im[0]), float(im[1]), pdX, pdc, a.mem))
    return dX, dc
