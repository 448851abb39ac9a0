"""Synthetic SMPL-topology constants and per-batch inputs (numpy only).

The real SMPL `model.pkl` and the LSP/UP datasets are not available offline, so
tests and `bench.py` use random arrays of the real shapes with the real 24-joint
parent tree (BASELINE.json north_star; recipe in SURVEY.md §8d).  The model dict
uses exactly the pickle keys the reference loader reads
(reference src/tf_smpl/batch_smpl.py:31-79), so `write_pkl()` output is consumed
unchanged by the reference's `SMPL(pkl_path)` and by ours.
"""
import pickle

import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_BASIS = 207
NUM_KEYPOINTS = 19

# kintree_table[0] of the real SMPL model (root's parent is uint32(-1)).
SMPL_PARENTS_U32 = np.array(
    [4294967295, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
    dtype=np.uint32)


def _sparse_regressor(rng, rows, num_verts, nnz):
    import scipy.sparse as sp
    nnz = min(nnz, num_verts)
    r, c, v = [], [], []
    for j in range(rows):
        cols = rng.choice(num_verts, size=nnz, replace=False)
        w = rng.dirichlet(np.ones(nnz))
        r += [j] * nnz
        c += list(cols)
        v += list(w)
    v = np.array(v).astype(np.float32).astype(np.float64)
    return sp.csc_matrix((v, (np.array(r), np.array(c))), shape=(rows, num_verts))


def make_model(seed=0, num_verts=NUM_VERTS, weights_profile="dense", regressor_nnz=32):
    """Dict with the reference pickle's keys and dtypes (float64 arrays, scipy
    sparse regressors, uint32 kintree_table).  Every value is exactly
    representable in float32, so an fp64 oracle and the fp32 kernels start from
    bit-identical constants."""
    rng = np.random.default_rng(seed)
    V = num_verts
    m = {}
    m["v_template"] = rng.normal(size=(V, 3)) * np.array([0.25, 0.45, 0.10])
    m["shapedirs"] = rng.normal(size=(V, 3, NUM_BETAS)) * 0.01
    m["posedirs"] = rng.normal(size=(V, 3, NUM_POSE_BASIS)) * 0.003
    m["J_regressor"] = _sparse_regressor(rng, NUM_JOINTS, V, regressor_nnz)
    m["cocoplus_regressor"] = _sparse_regressor(rng, NUM_KEYPOINTS, V, regressor_nnz)
    if weights_profile == "dense":
        w = rng.dirichlet(np.full(NUM_JOINTS, 0.2), size=V)
    elif weights_profile == "smpl_like":
        parents = SMPL_PARENTS_U32.astype(np.int64)
        parents[0] = 0
        w = np.zeros((V, NUM_JOINTS))
        main = rng.integers(0, NUM_JOINTS, size=V)
        for v in range(V):
            j = int(main[v])
            nb = {j, int(parents[j])}
            kids = [c for c in range(1, NUM_JOINTS) if parents[c] == j]
            cand = kids + [int(parents[int(parents[j])])] + list(range(NUM_JOINTS))
            for c in cand:
                if len(nb) >= 4:
                    break
                nb.add(int(c))
            nb = sorted(nb)
            w[v, nb] = rng.dirichlet(np.ones(len(nb)))
    else:
        raise ValueError("weights_profile must be 'dense' or 'smpl_like'")
    m["weights"] = w
    for k in ("v_template", "shapedirs", "posedirs", "weights"):
        m[k] = m[k].astype(np.float32).astype(np.float64)
    kt = np.zeros((2, NUM_JOINTS), dtype=np.uint32)
    kt[0] = SMPL_PARENTS_U32
    kt[1] = np.arange(NUM_JOINTS, dtype=np.uint32)
    m["kintree_table"] = kt
    return m


def write_pkl(model, path):
    with open(path, "wb") as f:
        pickle.dump(model, f, protocol=2)


def make_inputs(batch, seed=1000, num_keypoints=NUM_KEYPOINTS, dtype=np.float32):
    """beta/theta/cam/kp_gt as SURVEY.md §8d: sample 0 has theta == 0 (epsilon
    path of batch_rodrigues), sample 1 has every keypoint invisible."""
    rng = np.random.default_rng(seed)
    B, K = batch, num_keypoints
    beta = np.clip(rng.normal(size=(B, NUM_BETAS)), -3, 3)
    theta = rng.normal(size=(B, NUM_JOINTS, 3)) * 0.3
    theta[:, 0, :] = np.array([np.pi, 0, 0]) + rng.normal(size=(B, 3)) * 0.2
    theta = theta.reshape(B, 72)
    theta[0] = 0.0
    cam = np.stack([rng.uniform(0.6, 1.1, size=B), rng.normal(size=B) * 0.1, rng.normal(size=B) * 0.1], axis=1)
    xy = rng.uniform(-1, 1, size=(B, K, 2))
    p_vis = np.full(K, 0.7)
    p_vis[14:] = 0.35
    vis = (rng.uniform(size=(B, K)) < p_vis).astype(np.float64)
    if B > 1:
        vis[1] = 0.0
    kp_gt = np.concatenate([xy * vis[:, :, None], vis[:, :, None]], axis=2)  # data_loader.py:206-207 zeroes invisible rows
    return {"beta": beta.astype(dtype), "theta": theta.astype(dtype), "cam": cam.astype(dtype),
            "kp_gt": kp_gt.astype(dtype)}


def make_silhouettes(batch, seed=2000, img_size=224, a_range=(25, 45), b_range=(60, 95)):
    """seg [B,H,W,1] float32 in {0,1}: filled axis-aligned ellipses; sample 2 is
    empty (undefined in the reference; defined here to contribute 0)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:img_size, 0:img_size]
    seg = np.zeros((batch, img_size, img_size, 1), dtype=np.float32)
    for i in range(batch):
        cx, cy = img_size / 2 + rng.uniform(-15, 15, size=2)
        a = rng.uniform(*a_range)
        b = rng.uniform(*b_range)
        if i == 2:
            continue
        seg[i, :, :, 0] = (((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1.0)
    return seg


def silhouette_points(seg):
    """`tf.cast(tf.where(seg > 0)[:, :3], float32)` of reference
    src/trainer.py:443 -> rows (n, row, col), row-major order."""
    idx = np.argwhere(seg > 0)[:, :3]
    return idx.astype(np.float32)


def make_gp_inputs(m_samples, seed=3000, dtype=np.float32):
    """The four critic-input gradients of reference src/trainer.py:566-572:
    [M,13,13], [M,14,3], [M,10], [M,23,3,3]."""
    rng = np.random.default_rng(seed)
    shapes = [(13, 13), (14, 3), (10,), (23, 3, 3)]
    return [(rng.normal(size=(m_samples,) + s) / np.sqrt(np.prod(s)) + 0.05).astype(dtype) for s in shapes]
