"""Generate tests/golden/*.npz from the reference's own code (Oracle A).
TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The reference has no fixtures of its own, so these files ARE the pin: outputs
of src/tf_smpl/batch_smpl.py, batch_lbs.py, projection.py and src/ops.py
executed unchanged (fp64, torch-CPU `tensorflow` shim) on seeded synthetic
inputs, with gradients from autograd through that code.

  smpl_small.npz  V=160 model stored in full + B=6 inputs + every output.
  smpl_full.npz   V=6890 model regenerated from its seed (sha256 of the fp32
                  constants stored) + B=8 inputs + outputs (verts sub-sampled).
"""
import hashlib
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hpe_b200  # noqa: E402,F401
from hpe_b200 import synthetic  # noqa: E402
from oracle.run_reference import Reference  # noqa: E402

IMG = 224.0
VERT_STRIDE = 53


def model_digest(model):
    h = hashlib.sha256()
    for k in ("v_template", "shapedirs", "posedirs", "weights"):
        h.update(np.ascontiguousarray(model[k], dtype=np.float32).tobytes())
    for k in ("J_regressor", "cocoplus_regressor"):
        h.update(np.ascontiguousarray(np.asarray(model[k].todense()), dtype=np.float32).tobytes())
    h.update(np.ascontiguousarray(model["kintree_table"]).tobytes())
    return h.hexdigest()


def compact_points(pts, keep):
    """Renumber image ids so the reference (which cannot take an empty image)
    sees only the non-empty ones."""
    out = []
    for n, i in enumerate(keep):
        rows = pts[pts[:, 0] == i].copy()
        rows[:, 0] = n
        out.append(rows)
    return np.concatenate(out, axis=0)


def run_case(ref, model, B, seed, sil_kwargs, joint_type="cocoplus"):
    tf, torch = ref.tf, ref.torch
    smpl = ref.load_smpl(model, joint_type)
    K = 19 if joint_type == "cocoplus" else 14
    inp = synthetic.make_inputs(B, seed=seed, num_keypoints=K, dtype=np.float32)
    seg = synthetic.make_silhouettes(B, seed=seed + 1, **sil_kwargs)
    pts = synthetic.silhouette_points(seg)
    beta = ref.tensor(inp["beta"].astype(np.float64), True)
    theta = ref.tensor(inp["theta"].astype(np.float64), True)
    cam = ref.tensor(inp["cam"].astype(np.float64), True)
    kp_gt = ref.tensor(inp["kp_gt"].astype(np.float64))
    verts, joints, Rs = smpl(beta, theta, get_skin=True)
    Jtr = smpl.J_transformed
    kp = ref.projection.batch_orth_proj_idrot(joints, cam)
    kp_loss = ref.ops.kp_reprojection_loss(kp_gt, kp)
    sil_pred = ref.projection.reproject_vertices(verts, cam, tf.constant([IMG, IMG], ref.dtype))
    keep = [i for i in range(B) if np.any(pts[:, 0] == i)]
    pts_k = compact_points(pts, keep)
    mesh_loss = ref.ops.mesh_reprojection_loss(ref.tensor(pts_k.astype(np.float64)), sil_pred[keep], len(keep))
    out = dict(inp)
    out.update(seg_points=pts, verts=ref.np(verts), joints=ref.np(joints), Rs=ref.np(Rs), J_transformed=ref.np(Jtr),
               kp_pred=ref.np(kp), kp_loss=float(kp_loss), sil_pred=ref.np(sil_pred), mesh_loss=float(mesh_loss))
    # gradients of the kp loss alone (BASELINE config 2)
    g = torch.autograd.grad(kp_loss, [beta, theta, cam], retain_graph=True)
    out.update(kp_d_beta=g[0].numpy(), kp_d_theta=g[1].numpy(), kp_d_cam=g[2].numpy())
    # trainer weighting (src/config.py:67-68, src/trainer.py:433,448): 60*kp + 0.001*mesh
    total = 60.0 * kp_loss + 0.001 * mesh_loss
    g = torch.autograd.grad(total, [beta, theta, cam, sil_pred], retain_graph=True)
    out.update(step_d_beta=g[0].numpy(), step_d_theta=g[1].numpy(), step_d_cam=g[2].numpy())
    gsp = torch.autograd.grad(mesh_loss, [sil_pred], retain_graph=True)[0].numpy()
    out.update(mesh_d_sil_pred=gsp)
    # arbitrary upstream on all three outputs
    rng = np.random.default_rng(seed + 7)
    up_v = rng.normal(size=tuple(verts.shape)).astype(np.float32)
    up_j = rng.normal(size=tuple(joints.shape)).astype(np.float32)
    up_R = rng.normal(size=tuple(Rs.shape)).astype(np.float32)
    tot = (verts * ref.tensor(up_v.astype(np.float64))).sum() + (joints * ref.tensor(up_j.astype(np.float64))).sum() \
        + (Rs * ref.tensor(up_R.astype(np.float64))).sum()
    g = torch.autograd.grad(tot, [beta, theta])
    out.update(up_verts=up_v, up_joints=up_j, up_Rs=up_R, up_d_beta=g[0].numpy(), up_d_theta=g[1].numpy())
    # gradient penalty (src/ops.py:153) and its gradient
    gp_in = synthetic.make_gp_inputs(3 * B, seed=seed + 11, dtype=np.float32)
    tin = [ref.tensor(x.astype(np.float64), True) for x in gp_in]
    pen = ref.ops.compute_gradient_penalty(tin)
    gg = torch.autograd.grad(pen, tin)
    for n, (x, d) in enumerate(zip(gp_in, gg)):
        out["gp_in%d" % n] = x
        out["gp_grad%d" % n] = d.numpy()
    out["gp_penalty"] = float(pen)
    return out


def main():
    warnings.filterwarnings("ignore")
    ref = Reference(float64=True)
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)

    small = synthetic.make_model(seed=3, num_verts=160, regressor_nnz=12)
    res = run_case(ref, small, B=6, seed=100, sil_kwargs=dict(a_range=(4, 8), b_range=(7, 13)))
    consts = dict(
        m_v_template=small["v_template"].astype(np.float32), m_shapedirs=small["shapedirs"].astype(np.float32),
        m_posedirs=small["posedirs"].astype(np.float32), m_weights=small["weights"].astype(np.float32),
        m_J_regressor=np.asarray(small["J_regressor"].todense()).astype(np.float32),
        m_cocoplus_regressor=np.asarray(small["cocoplus_regressor"].todense()).astype(np.float32),
        m_kintree_table=small["kintree_table"])
    np.savez_compressed(os.path.join(gdir, "smpl_small.npz"), **consts, **res)

    lsp = run_case(ref, small, B=3, seed=300, sil_kwargs=dict(a_range=(4, 8), b_range=(7, 13)), joint_type="lsp")
    keep = ("beta", "theta", "cam", "kp_gt", "joints", "kp_pred", "kp_loss", "kp_d_beta", "kp_d_theta", "kp_d_cam")
    np.savez_compressed(os.path.join(gdir, "smpl_small_lsp.npz"), **{k: lsp[k] for k in keep})

    full = synthetic.make_model(seed=0)
    res = run_case(ref, full, B=8, seed=200, sil_kwargs=dict(a_range=(6, 10), b_range=(10, 16)))
    for k in ("verts", "sil_pred", "mesh_d_sil_pred"):
        res[k + "_sub"] = res.pop(k)[:, ::VERT_STRIDE]
    res["verts_sum"] = float(np.sum(np.abs(res["verts_sub"])))
    res["model_seed"] = 0
    res["model_sha256"] = model_digest(full)
    res["vert_stride"] = VERT_STRIDE
    np.savez_compressed(os.path.join(gdir, "smpl_full.npz"), **res)
    for f in sorted(os.listdir(gdir)):
        print(f, os.path.getsize(os.path.join(gdir, f)))


if __name__ == "__main__":
    main()
