"""human-pose-estimation_b200 -- B200-native SMPL + reprojection-loss hot path.

Host side of libsmplb.so (include/smplb.h), mirroring the reference's Python
interface for this path (maxpit/human-pose-estimation):

    reference module                      here
    src/tf_smpl/batch_smpl.py   SMPL      hpe_b200.tf_smpl.batch_smpl.SMPL
    src/tf_smpl/batch_lbs.py              hpe_b200.tf_smpl.batch_lbs
    src/tf_smpl/projection.py             hpe_b200.tf_smpl.projection
    src/ops.py                            hpe_b200.ops

Arrays are numpy (host) or runtime.DeviceArray (device); results come back in
the same kind.  Gradients, which the reference gets from TF autodiff, are
explicit: SMPL.backward, SMPL.step, *_backward.  Import as `import hpe_b200`
(alias module at the repo root).  No torch, no CPU fallback.
"""
from . import synthetic  # noqa: F401
from ._lib import DEVICE, HOST, LIB_PATH, SmplbError  # noqa: F401

__all__ = ["synthetic", "SmplbError", "HOST", "DEVICE", "LIB_PATH"]
