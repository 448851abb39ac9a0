"""Loss functions on libsmplb.so, mirroring reference src/ops.py:
kp_reprojection_loss (:35), find_nearest_neighbors (:60), bidirectional_dist
(:83), mesh_reprojection_loss (:117), compute_gradient_penalty (:153)."""
import ctypes as C

import numpy as np

from . import runtime, synthetic
from ._lib import GP_FLOATS, check, lib

_DEFAULT_CTX = {}


def _ctx_for(x, parents=None, device=0):
    """Context for a model-free op: the DeviceArray's own context, else a
    per-(device, parents) context over a 1-vertex dummy model."""
    if isinstance(x, runtime.DeviceArray) and parents is None:
        return x.ctx
    if parents is None:
        parents = synthetic.SMPL_PARENTS_U32.astype(np.int32)
    parents = np.asarray(parents).astype(np.int32)
    key = (device if not isinstance(x, runtime.DeviceArray) else x.ctx.device, tuple(int(p) for p in parents))
    if key not in _DEFAULT_CTX:
        z = np.zeros
        _DEFAULT_CTX[key] = runtime.Context(z((1, 3)), z((10, 3)), z((207, 3)), z((1, 24)), z((1, 24)), z((1, 1)),
                                            parents, device=key[0], max_batch=1)
    return _DEFAULT_CTX[key]


def kp_reprojection_loss_parts(kp_gt, kp_pred, want_grad=False):
    """(abs_sum, num_present[, d_kp_pred unscaled]) of kp_reprojection_loss:
    the numerator and the integer count, kept apart so shards can be
    all-reduced exactly."""
    ctx = _ctx_for(kp_pred)
    a = runtime.Args(ctx)
    K = int(kp_gt.shape[-2])
    N = int(np.prod(kp_gt.shape[:-2])) if len(kp_gt.shape) > 2 else 1
    pg, pp = a.inp(kp_gt, (N, K, 3)), a.inp(kp_pred, (N, K, 2))
    s, ps = a.out((1,))
    n, pn = a.out((1,), dtype=np.int64)
    d, pd = a.out((N, K, 2), want=want_grad)
    check(lib().smplb_kp_loss(ctx.handle, N, K, pg, pp, ps, pn, pd, a.mem))
    if isinstance(s, runtime.DeviceArray):
        s, n = s.numpy(), n.numpy()
    res = (float(s[0]), int(n[0]))
    return res + (d,) if want_grad else res


def kp_reprojection_loss(kp_gt, kp_pred, scale=1., name="kp_reprojection_loss"):
    """sum(vis * |kp_gt - kp_pred|) / (2 * #visible), 0 if nothing is visible.
    kp_gt N x K x 3 (x, y, vis), kp_pred N x K x 2.  `scale` is unused, as in
    the reference."""
    s, n = kp_reprojection_loss_parts(kp_gt, kp_pred)
    return np.float32(s / n) if n > 0 else np.float32(0.0)


def silhouette_csr(silhouette_gt, batch_size):
    """[P,3] rows (n, row, col) -> (points_xy [P,2] float32 with (x,y) =
    (col,row), offsets [batch+1] int32), rows of image i kept in their
    original order (ops.py:123-125 selects col0 == i and stacks (col2, col1))."""
    sg = np.asarray(silhouette_gt)
    n = sg[:, 0].astype(np.int64)
    keep = (n >= 0) & (n < batch_size) & (sg[:, 0] == n)
    order = np.argsort(n[keep], kind="stable")
    rows = sg[keep][order]
    counts = np.bincount(n[keep], minlength=batch_size)
    offsets = np.zeros(batch_size + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(counts)
    pts = np.ascontiguousarray(np.stack([rows[:, 2], rows[:, 1]], axis=1), dtype=np.float32)
    return pts, offsets


def silhouette_csr_device(seg, cap=None):
    """On-device version of silhouette_csr for a dense mask: seg [B,H,W] or [B,H,W,1] (numpy or
    DeviceArray) -> (points_xy [cap,2], offsets [B+1]) of the same kind.  `cap` bounds the
    number of points kept (default B*H*W)."""
    ctx = _ctx_for(seg)
    a = runtime.Args(ctx)
    shp = tuple(seg.shape)
    B, H, W = int(shp[0]), int(shp[1]), int(shp[2])
    cap = B * H * W if cap is None else int(cap)
    ps = a.inp(seg, (B, H, W))
    pts, pp = a.out((cap, 2))
    offs, po = a.out((B + 1,), dtype=np.int32)
    check(lib().smplb_silhouette_csr(ctx.handle, B, H, W, ps, pp, cap, po, a.mem))
    return pts, offs


def _mesh_call(ctx, pts, offsets, sil_pred, want_grad, want_idx):
    a = runtime.Args(ctx)
    N, V = int(sil_pred.shape[0]), int(sil_pred.shape[1])
    P = int(pts.shape[0])
    ps = a.inp(sil_pred, (N, V, 2))
    if a.mem == runtime.DEVICE:
        if not isinstance(pts, runtime.DeviceArray):
            pts = ctx.to_device(pts) if P > 0 else None
            offsets = ctx.to_device(offsets, np.int32)
        pp = pts.ptr if P > 0 else None
        po = offsets.ptr
    else:
        pp = a.inp(pts, (P, 2)) if P > 0 else None
        po = a.inp(offsets, (N + 1,), dtype=np.int32)
    loss, pl = a.out((1,))
    g, pg = a.out((N, V, 2), want=want_grad)
    iab, pia = a.out((max(P, 1),), dtype=np.int32, want=want_idx)
    iba, pib = a.out((N, V), dtype=np.int32, want=want_idx)
    check(lib().smplb_mesh_reproj_loss(ctx.handle, N, V, pp, po, P, ps, pl, pg, pia, pib, a.mem))
    return loss, g, iab, iba


def find_nearest_neighbors(A, B):
    """A num_A x 2, B num_B x 2 -> (ind_AB [num_A], ind_BA [num_B]): nearest
    neighbour of every A in B and of every B in A under the fp32 expansion
    -2AB^T + |A|^2 + |B|^2, first index on ties."""
    ctx = _ctx_for(B)
    A_h = A.numpy() if isinstance(A, runtime.DeviceArray) else np.ascontiguousarray(A, dtype=np.float32)
    nA = A_h.shape[0]
    offs = np.array([0, nA], dtype=np.int32)
    Bb = B if not isinstance(B, runtime.DeviceArray) else B
    if isinstance(B, runtime.DeviceArray):
        nB = B.shape[0]
        view = runtime.DeviceArray.__new__(runtime.DeviceArray)
        view.ctx, view.shape, view.dtype, view.nbytes, view.ptr = B.ctx, (1, nB, 2), B.dtype, B.nbytes, B.ptr
        view.free = lambda: None
        Bb = view
    else:
        Bb = np.ascontiguousarray(B, dtype=np.float32)[None]
    _, _, iab, iba = _mesh_call(ctx, A_h, offs, Bb, False, True)
    if isinstance(iab, runtime.DeviceArray):
        iab, iba = iab.numpy(), iba.numpy()
    return iab[:nA].astype(np.int64), iba.reshape(-1).astype(np.int64)


def bidirectional_dist(A, B):
    """sum_b ||B_b - A[nn(b)]||_2 + sum_a ||A_a - B[nn(a)]||_1 (ops.py:83-102)."""
    ctx = _ctx_for(B)
    A_h = A.numpy() if isinstance(A, runtime.DeviceArray) else np.ascontiguousarray(A, dtype=np.float32)
    B_h = B.numpy() if isinstance(B, runtime.DeviceArray) else np.ascontiguousarray(B, dtype=np.float32)
    offs = np.array([0, A_h.shape[0]], dtype=np.int32)
    loss, _, _, _ = _mesh_call(ctx, A_h, offs, B_h[None], False, False)
    # the C ABI returns dist / (3 + num_B) (mesh_reprojection_loss' scaling, ops.py:129-130)
    return np.float32(loss[0] * (3 + B_h.shape[0]))


def mesh_reprojection_loss(silhouette_gt, silhouette_pred, batch_size, name="mesh_reprojection_loss",
                           want_grad=False):
    """silhouette_gt [sum P_i, 3] rows (image, row, col); silhouette_pred
    N x 6890 x 2 pixel coordinates -> sum_i bidirectional_dist_i / (3 + 6890).
    With want_grad also returns d loss / d silhouette_pred."""
    ctx = _ctx_for(silhouette_pred)
    sg = silhouette_gt.numpy() if isinstance(silhouette_gt, runtime.DeviceArray) else silhouette_gt
    pts, offs = silhouette_csr(sg, int(batch_size))
    loss, g, _, _ = _mesh_call(ctx, pts, offs, silhouette_pred, want_grad, False)
    lv = loss.numpy()[0] if isinstance(loss, runtime.DeviceArray) else loss[0]
    return (np.float32(lv), g) if want_grad else np.float32(lv)


def compute_gradient_penalty(gradients, debug=False, want_sums=False):
    """gradients: [M,13,13], [M,14,3], [M,10], [M,23,3,3] -> sum_i (1 -
    ||mean_axis0 g_i||)^2 (ops.py:153-172).  With want_sums also returns the
    428 column sums a multi-GPU caller all-reduces."""
    g0, g1, g2, g3 = gradients
    ctx = _ctx_for(g0)
    a = runtime.Args(ctx)
    M = int(g0.shape[0])
    p = [a.inp(g0, (M, 169)), a.inp(g1, (M, 42)), a.inp(g2, (M, 10)), a.inp(g3, (M, 207))]
    pen, pp = a.out((1,))
    sums, psum = a.out((GP_FLOATS,), want=want_sums)
    check(lib().smplb_gradient_penalty(ctx.handle, M, p[0], p[1], p[2], p[3], pp, psum, a.mem))
    pv = pen.numpy()[0] if isinstance(pen, runtime.DeviceArray) else pen[0]
    if debug:
        print("penalty", float(pv))
    return (np.float32(pv), sums) if want_sums else np.float32(pv)


def gradient_penalty_step(gradients, out=None, m_total=None):
    """Device-resident penalty + its gradient in one asynchronous call pair (what the critic stage of
    src/trainer.py:566-578 needs every step): gradients = four DeviceArrays; `out` keeps the preallocated
    outputs {"penalty" [1], "col_sums" [428], "d0".."d3"} across calls.  No host synchronisation."""
    g0, g1, g2, g3 = gradients
    ctx = _ctx_for(g0)
    M = int(g0.shape[0])
    out = {} if out is None else out
    if "penalty" not in out:
        out["penalty"] = ctx.empty((1,))
        out["col_sums"] = ctx.empty((GP_FLOATS,))
        for i, shp in enumerate([(M, 13, 13), (M, 14, 3), (M, 10), (M, 23, 3, 3)]):
            out["d%d" % i] = ctx.empty(shp)
    check(lib().smplb_gradient_penalty(ctx.handle, M, g0.ptr, g1.ptr, g2.ptr, g3.ptr, out["penalty"].ptr, out["col_sums"].ptr,
                                       runtime.DEVICE))
    check(lib().smplb_gradient_penalty_backward(ctx.handle, M, int(m_total or M), out["col_sums"].ptr, out["d0"].ptr,
                                                out["d1"].ptr, out["d2"].ptr, out["d3"].ptr, runtime.DEVICE))
    return out


def gradient_penalty_from_sums(col_sums, m_total):
    ctx = _ctx_for(col_sums)
    a = runtime.Args(ctx)
    ps = a.inp(col_sums, (GP_FLOATS,))
    pen, pp = a.out((1,))
    check(lib().smplb_gradient_penalty_from_sums(ctx.handle, int(m_total), ps, pp, a.mem))
    return np.float32(pen.numpy()[0] if isinstance(pen, runtime.DeviceArray) else pen[0])


def gradient_penalty_backward(col_sums, m_local, m_total=None):
    """d penalty / d gradients, as four arrays shaped like the inputs."""
    ctx = _ctx_for(col_sums)
    a = runtime.Args(ctx)
    ps = a.inp(col_sums, (GP_FLOATS,))
    M = int(m_local)
    shapes = [(M, 13, 13), (M, 14, 3), (M, 10), (M, 23, 3, 3)]
    outs = [a.out(s) for s in shapes]
    check(lib().smplb_gradient_penalty_backward(ctx.handle, M, int(m_total or M), ps, outs[0][1], outs[1][1],
                                                outs[2][1], outs[3][1], a.mem))
    return [o[0] for o in outs]
