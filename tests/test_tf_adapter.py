"""The `tf.custom_gradient` adapter (reference call site src/trainer.py:411)
executed under oracle/tf_shim: gradients that flow through the adapter into (beta, theta) must equal autograd
through the line-by-line torch restatement of the reference (oracle/smpl_torch.py) for a loss that touches all
three outputs, for the keypoint loss alone (verts / Rs receive zero upstream gradients), and when several
forwards are taken before the tape is differentiated (num_stage forwards, trainer.py:391-411)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err
from hpe_b200 import synthetic
from oracle import smpl_numpy as onp


@pytest.fixture()
def tf():
    shim = os.path.join(ROOT, "oracle", "tf_shim")
    saved = sys.modules.pop("tensorflow", None)
    sys.path.insert(0, shim)
    import tensorflow as tf_shim
    import torch
    tf_shim.set_float(torch.float32)
    yield tf_shim
    sys.path.remove(shim)
    sys.modules.pop("tensorflow", None)
    if saved is not None:
        sys.modules["tensorflow"] = saved


class OracleSMPL(object):
    """Stand-in with the facade's interface (SMPL.__call__ / .backward, depth-1 tape) on the numpy oracle."""

    def __init__(self, model):
        self.o = onp.SMPL(model, dtype=np.float64)
        self.size = self.o.size
        self.num_keypoints = self.o.joint_regressor.shape[1]
        self.calls = 0

    def __call__(self, beta, theta, get_skin=False):
        self.calls += 1
        self.saved = (np.asarray(beta, np.float64), np.asarray(theta, np.float64))
        v, j, R = self.o(self.saved[0], self.saved[1], get_skin=True)
        return v.astype(np.float32), j.astype(np.float32), R.astype(np.float32)

    def backward(self, d_verts=None, d_joints=None, d_Rs=None, batch=None):
        f = lambda x: None if x is None else np.asarray(x, np.float64)   # noqa: E731
        db, dt = onp.smpl_backward(self.o, self.saved[0], self.saved[1], f(d_verts), f(d_joints), f(d_Rs))
        return db.astype(np.float32), dt.astype(np.float32)


def _check_adapter(tf, smpl, model, tol):
    import torch
    from hpe_b200.tf_adapter import make_tf_smpl
    from oracle import smpl_torch as ot
    smpl_tf = make_tf_smpl(smpl)
    ts = ot.SMPL(model, dtype=torch.float64)
    K = smpl.num_keypoints
    inp = [synthetic.make_inputs(4, seed=40 + i, num_keypoints=K) for i in range(2)]
    rng = np.random.default_rng(0)
    V = smpl.size[0]
    wv, wj, wR = rng.normal(size=(4, V, 3)), rng.normal(size=(4, K, 3)), rng.normal(size=(4, 24, 3, 3))

    def losses(fn, beta, theta, cam, kp_gt, f64):
        dt = torch.float64 if f64 else torch.float32
        verts, joints, Rs = fn(beta, theta)
        full = (verts * torch.as_tensor(wv, dtype=dt)).sum() + (joints * torch.as_tensor(wj, dtype=dt)).sum() \
            + (Rs * torch.as_tensor(wR, dtype=dt)).sum()
        kp = ot.batch_orth_proj_idrot(joints, cam)
        return full, ot.kp_reprojection_loss(kp_gt, kp)

    def leaf(a, f64):
        return torch.tensor(a, dtype=torch.float64 if f64 else torch.float32, requires_grad=True)

    want, got = [], []
    for which in (0, 1):                              # 0: every output used; 1: keypoint loss only
        x = inp[which]
        b64, t64, c64 = leaf(x["beta"], True), leaf(x["theta"], True), leaf(x["cam"], True)
        l = losses(lambda b, t: ts(b, t, get_skin=True), b64, t64, c64, torch.tensor(x["kp_gt"], dtype=torch.float64), True)[which]
        want.append(torch.autograd.grad(l, [b64, t64]))
    # the adapter: BOTH forwards first (as the stages of train_step), then the gradients in reverse order
    leaves, ls = [], []
    for which in (0, 1):
        x = inp[which]
        b, t, c = leaf(x["beta"], False), leaf(x["theta"], False), leaf(x["cam"], False)
        leaves.append((b, t))
        ls.append(losses(lambda bb, tt: smpl_tf(tf.convert_to_tensor(bb), tf.convert_to_tensor(tt)), b, t, c,
                         torch.tensor(x["kp_gt"]), False)[which])
    for which in (1, 0):
        got.insert(0, torch.autograd.grad(ls[which], list(leaves[which])))
    for which in (0, 1):
        for g, w in zip(got[which], want[which]):
            assert rel_err(g.numpy(), w.numpy()) < tol, which
    # stage 1's backward found its own forward state; stage 0's had to be restored exactly once
    assert smpl_tf.state["reruns"] == 1


def test_tf_adapter_under_shim_cpu(tf, small_model):
    _check_adapter(tf, OracleSMPL(small_model), small_model, 1e-5)


@pytest.mark.gpu
def test_tf_adapter_under_shim_gpu(tf, small_model):
    from hpe_b200.tf_smpl.batch_smpl import SMPL
    _check_adapter(tf, SMPL(small_model, max_batch=8), small_model, 1e-4)
