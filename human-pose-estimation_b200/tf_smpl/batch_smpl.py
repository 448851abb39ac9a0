"""SMPL(beta, theta) -> (verts, joints, Rs) on libsmplb.so.

Mirrors reference src/tf_smpl/batch_smpl.py:25-160: same constructor arguments,
same public attributes, same call signature and return convention, including
the `J_transformed` attribute set as a side effect of a call (:135).
"""
import pickle

import numpy as np

from .. import runtime
from .. import _lib
from .._lib import check, lib


def undo_chumpy(x):
    # batch_smpl.py:21-22: chumpy objects expose their value as .r
    return x if isinstance(x, np.ndarray) else x.r


def load_model_arrays(dd, joint_type="cocoplus"):
    """The constant re-layout SMPL.__init__ performs (batch_smpl.py:34-81), in
    numpy: returns fp32 arrays ready for smplb_create."""
    if joint_type not in ("cocoplus", "lsp"):
        # the reference prints 'BAD!!' and drops into ipdb (batch_smpl.py:83-86)
        raise ValueError('Unknown joint type: %s, it must be either "cocoplus" or "lsp"' % joint_type)
    v_template = np.asarray(undo_chumpy(dd["v_template"]), dtype=np.float32)
    num_betas = dd["shapedirs"].shape[-1]
    shapedirs = np.reshape(undo_chumpy(dd["shapedirs"]), [-1, num_betas]).T.astype(np.float32)
    J_regressor = np.asarray(dd["J_regressor"].T.todense(), dtype=np.float32)
    num_pose_basis = dd["posedirs"].shape[-1]
    posedirs = np.reshape(undo_chumpy(dd["posedirs"]), [-1, num_pose_basis]).T.astype(np.float32)
    parents = dd["kintree_table"][0].astype(np.int32)
    weights = np.asarray(undo_chumpy(dd["weights"]), dtype=np.float32)
    joint_regressor = np.asarray(dd["cocoplus_regressor"].T.todense(), dtype=np.float32)
    if joint_type == "lsp":
        joint_regressor = joint_regressor[:, :14]
    return dict(v_template=v_template, shapedirs=np.ascontiguousarray(shapedirs),
                posedirs=np.ascontiguousarray(posedirs), J_regressor=np.ascontiguousarray(J_regressor),
                weights=np.ascontiguousarray(weights), joint_regressor=np.ascontiguousarray(joint_regressor),
                parents=parents, num_betas=num_betas)


class SMPL(object):
    def __init__(self, pkl_path, joint_type="cocoplus", dtype=np.float32, device=0, max_batch=64):
        """pkl_path is the path to a SMPL model pickle (or the unpickled dict).
        `dtype` is accepted for signature parity; the kernels are fp32 like the
        reference's default.  `device`/`max_batch` size the GPU context."""
        if isinstance(pkl_path, dict):
            dd = pkl_path
        else:
            with open(pkl_path, "rb") as f:
                try:
                    dd = pickle.load(f)
                except UnicodeDecodeError:
                    f.seek(0)
                    dd = pickle.load(f, encoding="latin1")
        if np.dtype(dtype) != np.float32:
            raise ValueError("only float32 is supported (the reference's default dtype)")
        m = load_model_arrays(dd, joint_type)
        self.v_template = m["v_template"]
        self.size = [self.v_template.shape[0], 3]
        self.num_betas = m["num_betas"]
        self.shapedirs = m["shapedirs"]
        self.J_regressor = m["J_regressor"]
        self.posedirs = m["posedirs"]
        self.parents = m["parents"]
        self.weights = m["weights"]
        self.joint_regressor = m["joint_regressor"]
        self.J_transformed = None
        self.ctx = runtime.Context(self.v_template, self.shapedirs, self.posedirs, self.J_regressor, self.weights,
                                   self.joint_regressor, self.parents, device=device, max_batch=max_batch)
        self.num_keypoints = self.joint_regressor.shape[1]

    # -- forward -----------------------------------------------------------
    def __call__(self, beta, theta, get_skin=False, name=None):
        """beta: N x 10, theta: N x 72 (or 72 for N = 1, as data_loader.py:141
        calls it).  Returns joints (N x K x 3), or (verts, joints, Rs) if
        get_skin.  Updates self.J_transformed (N x 24 x 3)."""
        a = runtime.Args(self.ctx)
        N = int(beta.shape[0])
        pb = a.inp(beta, (N, self.num_betas))
        pt = a.inp(theta, (N, 72))
        V, K = self.size[0], self.num_keypoints
        verts, pv = a.out((N, V, 3), want=get_skin)
        joints, pj = a.out((N, K, 3))
        Rs, pR = a.out((N, 24, 3, 3), want=get_skin)
        Jtr, pJ = a.out((N, 24, 3))
        check(lib().smplb_smpl_forward(self.ctx.handle, N, pb, pt, pv, pj, pR, pJ, 0, a.mem))
        self.J_transformed = Jtr
        if get_skin:
            return verts, joints, Rs
        return joints

    def forward_joints_rotations(self, beta, theta):
        """joints and Rs without materialising verts (what the critic's "real" stream and the
        mocap preprocessing of data_loader.py:139-143 consume)."""
        a = runtime.Args(self.ctx)
        N = int(beta.shape[0])
        pb, pt = a.inp(beta, (N, self.num_betas)), a.inp(theta, (N, 72))
        joints, pj = a.out((N, self.num_keypoints, 3))
        Rs, pR = a.out((N, 24, 3, 3))
        check(lib().smplb_smpl_forward(self.ctx.handle, N, pb, pt, None, pj, pR, None, 0, a.mem))
        return joints, Rs

    def forward_into(self, beta, theta, n, verts=None, joints=None, Rs=None):
        """Device-resident serving call (src/predictor.py:141): the first `n` rows of the DeviceArrays
        beta / theta -> preallocated DeviceArrays (verts and Rs optional); asynchronous, no allocation."""
        ptr = lambda x: None if x is None else x.ptr   # noqa: E731
        check(lib().smplb_smpl_forward(self.ctx.handle, int(n), beta.ptr, theta.ptr, ptr(verts), joints.ptr, ptr(Rs), None, 0,
                                       runtime.DEVICE))

    # -- backward (TF autodiff in the reference, src/trainer.py:383,502) ----
    def backward(self, d_verts=None, d_joints=None, d_Rs=None, batch=None):
        """Gradients of the last call w.r.t. (beta, theta) for upstream gradients on
        verts / joints / Rs (each optional)."""
        a = runtime.Args(self.ctx)
        given = [x for x in (d_verts, d_joints, d_Rs) if x is not None]
        if not given:
            raise ValueError("at least one upstream gradient is required")
        N = int(given[0].shape[0]) if batch is None else batch
        V, K = self.size[0], self.num_keypoints
        pv = a.inp(d_verts, (N, V, 3))
        pj = a.inp(d_joints, (N, K, 3))
        pR = a.inp(d_Rs, (N, 24, 3, 3))
        d_beta, pb = a.out((N, self.num_betas))
        d_theta, pt = a.out((N, 72))
        check(lib().smplb_smpl_backward(self.ctx.handle, N, pv, pj, pR, pb, pt, a.mem))
        return d_beta, d_theta

    # -- one generator stage of Trainer.train_step + its backward -----------
    def step(self, beta, theta, cam, kp_gt, silhouette=None, w_kp=60.0, w_mesh=0.001, img_size=224.0,
             backward=True, want_verts=True, kp_count_override=0, out=None, skip=(), nowait=False, seg=None):
        """src/trainer.py:404-450 + :502 for one stage: SMPL forward, keypoint
        projection and loss, optional mesh-reprojection loss, and gradients
        w.r.t. beta/theta/cam.  `silhouette` = (points_xy [P,2], offsets [B+1])
        from ops.silhouette_csr, or `seg` = the dense mask [B,H,W(,1)] itself (compacted on the
        device inside the call, trainer.py:443).  Returns a dict.  `out` may hold preallocated
        outputs of the right kind to avoid allocations in a timed loop;
        want_verts="device" (host-buffer calls) computes verts but leaves them in
        device memory (out["verts_device_ptr"]) instead of copying 339 MB back; names in
        `skip` (of "joints", "Rs", "kp_pred") are not returned (host mode: not
        copied back).  nowait=True (host mode, pinned arrays only) returns without
        synchronising; call ctx.sync() before reading the outputs."""
        a = runtime.Args(self.ctx)
        N = int(beta.shape[0])
        V, K = self.size[0], self.num_keypoints
        pb, pt = a.inp(beta, (N, self.num_betas)), a.inp(theta, (N, 72))
        pc, pk = a.inp(cam, (N, 3)), a.inp(kp_gt, (N, K, 3))
        pp = po = None
        P = 0
        pseg = None
        if seg is not None:
            assert silhouette is None, "pass either silhouette= (CSR points) or seg= (dense mask)"
            H, W = int(seg.shape[1]), int(seg.shape[2])
            pseg = a.inp(seg, (N, H, W))
        if silhouette is not None:
            pts, offs = silhouette
            P = int(pts.shape[0])
            pp = a.inp(pts, (P, 2)) if P > 0 else None
            po = a.inp(offs, (N + 1,), dtype=np.int32)
        out = {} if out is None else out

        def o(name, shape, want=True):
            if name in skip:
                out[name] = None
                return None, None
            if name in out and out[name] is not None:
                x = out[name]
                return x, (x.ptr if isinstance(x, runtime.DeviceArray) else x.ctypes.data)
            x, p = a.out(shape, want=want)
            out[name] = x
            return x, p

        _, pv = o("verts", (N, V, 3), want_verts is True)
        _, pj = o("joints", (N, K, 3))
        _, pR = o("Rs", (N, 24, 3, 3))
        _, pkp = o("kp_pred", (N, K, 2))
        _, pl = o("loss_parts", (4,))
        _, pdb = o("d_beta", (N, self.num_betas), backward)
        _, pdt = o("d_theta", (N, 72), backward)
        _, pdc = o("d_cam", (N, 3), backward)
        flags = _lib.STEP_KEEP_VERTS if want_verts == "device" else 0
        mem = 2 if (nowait and a.mem == runtime.HOST) else a.mem
        if pseg is not None:
            check(lib().smplb_step_seg(self.ctx.handle, N, pb, pt, pc, pk, pseg, H, W, float(w_kp), float(w_mesh),
                                       float(img_size), int(kp_count_override), pv, pj, pR, pkp, pl, pdb, pdt, pdc, flags, mem))
        else:
            check(lib().smplb_step(self.ctx.handle, N, pb, pt, pc, pk, pp, po, P, float(w_kp), float(w_mesh),
                                   float(img_size), int(kp_count_override), pv, pj, pR, pkp, pl, pdb, pdt, pdc, flags, mem))
        if want_verts == "device":
            out["verts_device_ptr"] = self.last_verts_ptr()
        return out

    def last_verts_ptr(self):
        """Device address of the verts of the last call (workspace or caller buffer), or None."""
        import ctypes as C
        p = C.c_void_p()
        check(lib().smplb_last_verts(self.ctx.handle, C.byref(p)))
        return p.value
