"""The one piece of reference src/data_loader.py that runs the hot path: preprocess_poses
(:139-143), which the reference maps over the mocap dataset ONE pose at a time (SMPL at batch 1
inside tf.data, on the CPU).  Here the whole table goes through in one (or a few) large GPU
calls; verts are never materialised (joints come from the folded
keypoint path), so 3.9 M CMU poses are a few hundred milliseconds of GPU time."""
import numpy as np


def preprocess_poses(smpl, pose, shape, chunk=65536):
    """pose [N,72] (or [72]), shape [N,10] (or [10]) -> (joints [N,K,3], shape, rotations
    [N,24,3,3]), the tuple the reference's per-example map returns (with its leading batch-1
    axis folded into N)."""
    pose = np.asarray(pose, dtype=np.float32).reshape(-1, 72)
    shape = np.asarray(shape, dtype=np.float32).reshape(-1, pose.shape[0] and shape.size // pose.shape[0])
    N = pose.shape[0]
    joints = np.empty((N, smpl.num_keypoints, 3), np.float32)
    rots = np.empty((N, 24, 3, 3), np.float32)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        out = smpl.forward_joints_rotations(shape[s:e], pose[s:e])
        joints[s:e], rots[s:e] = out
    return joints, shape, rots
