"""TensorFlow-2 adapter: exposes the kernels to the reference's
`tf.GradientTape` code as differentiable ops, so `src/trainer.py` / `src/predictor.py` keep their
structure.  TensorFlow is NOT available in the build image (no wheel, no network), so this
module is import-guarded and untested here; it only composes calls that are tested
(SMPL.__call__, SMPL.backward) with `tf.custom_gradient` and `tf.numpy_function`.

    from hpe_b200.tf_adapter import make_tf_smpl
    smpl_tf = make_tf_smpl(SMPL("models/model.pkl"))
    verts, joints, Rs = smpl_tf(shapes, pose)          # inside the tape of trainer.py:383-411
"""


def make_tf_smpl(smpl):
    try:
        import tensorflow as tf
    except ImportError as e:   # pragma: no cover - TensorFlow is absent offline
        raise ImportError("hpe_b200.tf_adapter needs TensorFlow 2.x, which is not installed") from e
    import numpy as np

    def _fwd(beta, theta):
        verts, joints, Rs = smpl(np.asarray(beta), np.asarray(theta), get_skin=True)
        return verts, joints, Rs

    def _bwd(beta, theta, d_verts, d_joints, d_Rs):
        # the context keeps a depth-1 tape: re-run the forward of these inputs, then its backward
        smpl(np.asarray(beta), np.asarray(theta), get_skin=True)
        d_beta, d_theta = smpl.backward(np.asarray(d_verts), np.asarray(d_joints), np.asarray(d_Rs))
        return d_beta, d_theta

    @tf.custom_gradient
    def smpl_tf(beta, theta):
        verts, joints, Rs = tf.numpy_function(_fwd, [beta, theta], [tf.float32, tf.float32, tf.float32])
        n = beta.shape[0]
        verts.set_shape([n, smpl.size[0], 3])
        joints.set_shape([n, smpl.num_keypoints, 3])
        Rs.set_shape([n, 24, 3, 3])

        def grad(d_verts, d_joints, d_Rs):
            d_beta, d_theta = tf.numpy_function(_bwd, [beta, theta, d_verts, d_joints, d_Rs], [tf.float32, tf.float32])
            d_beta.set_shape(beta.shape)
            d_theta.set_shape(theta.shape)
            return d_beta, d_theta

        return (verts, joints, Rs), grad

    return smpl_tf
