"""Oracle A -- the reference's own hot-path files executed unchanged.
TEST INFRASTRUCTURE ONLY (never imported by the product).

Imports src/tf_smpl/batch_smpl.py, batch_lbs.py, projection.py and src/ops.py
straight from the read-only reference checkout with oracle/tf_shim's
torch-backed `tensorflow` stand-in on sys.path.  Works only where
/root/reference exists (the build container); on the GPU box the committed
golden vectors in tests/golden/ (written by oracle/make_golden.py from this
module) take its place.
"""
import importlib
import os
import sys
import tempfile

import numpy as np

REFERENCE_ROOT = os.environ.get("SMPLB_REFERENCE_ROOT", "/root/reference")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_shim")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "tf_smpl", "batch_smpl.py"))


class Reference(object):
    """Handles to the reference modules.  `float64=True` re-maps tf.float32 to
    torch.float64 so the very same code serves as a high-precision oracle."""

    def __init__(self, float64=True):
        if not available():
            raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
        import torch

        sys.dont_write_bytecode = True           # /root/reference is read-only
        for p in (_SHIM_DIR, REFERENCE_ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        import tensorflow as tf                   # the shim

        assert os.path.dirname(tf.__file__).startswith(_SHIM_DIR), "a real tensorflow shadowed the shim"
        self.torch = torch
        self.tf = tf
        self.dtype = torch.float64 if float64 else torch.float32
        tf.set_float(self.dtype)
        self.batch_smpl = importlib.import_module("src.tf_smpl.batch_smpl")
        self.batch_lbs = importlib.import_module("src.tf_smpl.batch_lbs")
        self.projection = importlib.import_module("src.tf_smpl.projection")
        self.ops = importlib.import_module("src.ops")

    def tensor(self, x, requires_grad=False):
        t = self.tf.convert_to_tensor(np.asarray(x), self.dtype if np.asarray(x).dtype.kind == "f" else None)
        if requires_grad:
            t = t.detach().clone().requires_grad_(True).as_subclass(self.tf.Tensor)
        return t

    def load_smpl(self, model, joint_type="cocoplus"):
        """Pickle `model` and let the reference's own loader read it."""
        import pickle

        self.tf.set_float(self.dtype)
        with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as f:
            pickle.dump(model, f, protocol=2)
            path = f.name
        try:
            return self.batch_smpl.SMPL(path, joint_type=joint_type, dtype=self.dtype)
        finally:
            os.unlink(path)

    @staticmethod
    def np(t):
        return t.detach().as_subclass(__import__("torch").Tensor).numpy()
