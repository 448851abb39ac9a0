"""TensorFlow-2 adapter: exposes the kernels to the reference's
`tf.GradientTape` code as one differentiable op, so `src/trainer.py:411` / `src/predictor.py:141`
keep their structure:

    from hpe_b200.tf_adapter import make_tf_smpl
    smpl_tf = make_tf_smpl(SMPL("models/model.pkl"))
    verts, joints, Rs = smpl_tf(shapes, pose)          # inside the tape of trainer.py:383-411

It composes SMPL.__call__ and SMPL.backward with `tf.custom_gradient` and `tf.numpy_function`.
TensorFlow itself is not installable in the build image; tests/test_tf_adapter.py exercises the
adapter against the test suite's torch-backed `tensorflow` stand-in (which implements those two
symbols on torch autograd) -- on the CPU with a numpy stand-in for the SMPL object, on the GPU
with the real one.

The context keeps a depth-1 tape (backward pairs with the last forward on it).  The adapter
numbers its forwards and only re-runs one when another forward came in between (several stages
of `train_step` call SMPL before the tape is differentiated: trainer.py:391-411 runs num_stage
forwards, then one backward through all of them in reverse).
"""


def make_tf_smpl(smpl):
    try:
        import tensorflow as tf
    except ImportError as e:   # pragma: no cover - TensorFlow is absent offline
        raise ImportError("hpe_b200.tf_adapter needs TensorFlow 2.x, which is not installed") from e
    import numpy as np

    state = {"seq": 0, "reruns": 0}

    def _fwd(beta, theta):
        state["seq"] += 1
        state["last"] = state["seq"]
        return smpl(np.asarray(beta, dtype=np.float32), np.asarray(theta, dtype=np.float32), get_skin=True)

    def _make_bwd(my_seq):
        def _bwd(beta, theta, d_verts, d_joints, d_Rs):
            if state.get("last") != my_seq:
                # another forward ran on the context since: restore this one's saved state
                smpl(np.asarray(beta, dtype=np.float32), np.asarray(theta, dtype=np.float32), get_skin=True)
                state["last"] = my_seq
                state["reruns"] += 1

            def up(x):   # an all-zero upstream gradient costs a 339 MB copy and a dense backward: drop it
                x = np.asarray(x, dtype=np.float32)
                return x if x.any() else None

            ups = [up(d_verts), up(d_joints), up(d_Rs)]
            if all(u is None for u in ups):
                return (np.zeros(np.shape(beta), np.float32), np.zeros(np.shape(theta), np.float32))
            d_beta, d_theta = smpl.backward(ups[0], ups[1], ups[2], batch=int(np.shape(beta)[0]))
            return np.asarray(d_beta, dtype=np.float32), np.asarray(d_theta, dtype=np.float32)
        return _bwd

    @tf.custom_gradient
    def smpl_tf(beta, theta):
        verts, joints, Rs = tf.numpy_function(_fwd, [beta, theta], [tf.float32, tf.float32, tf.float32])
        my_seq = state["seq"]
        n = beta.shape[0]
        verts.set_shape([n, smpl.size[0], 3])
        joints.set_shape([n, smpl.num_keypoints, 3])
        Rs.set_shape([n, 24, 3, 3])

        def grad(d_verts, d_joints, d_Rs):
            d_beta, d_theta = tf.numpy_function(_make_bwd(my_seq), [beta, theta, d_verts, d_joints, d_Rs],
                                                [tf.float32, tf.float32])
            d_beta.set_shape(beta.shape)
            d_theta.set_shape(theta.shape)
            return d_beta, d_theta

        return (verts, joints, Rs), grad

    smpl_tf.state = state
    return smpl_tf
