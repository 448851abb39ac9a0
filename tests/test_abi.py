"""CPU tests of the boundary: the C-ABI library loads, exports every symbol
include/smplb.h declares, fails loudly without a GPU, and the host-side
helpers behave like the reference's.  No compute calls are made."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from hpe_b200 import _lib, ops, synthetic


def header_functions():
    src = open(os.path.join(ROOT, "include", "smplb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smplb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libsmplb.so does not export %s" % n
    # the ctypes table binds exactly the header's functions (minus the two untyped getters)
    assert set(_lib.SIGNATURES) | {"smplb_last_error", "smplb_version"} == set(names)
    # test / tuning hooks are exported but live in the private header, not in the public ABI
    priv = open(os.path.join(ROOT, "human-pose-estimation_b200", "csrc", "smplb_debug.h")).read()
    for n in _lib.PRIVATE_SIGNATURES:
        assert n not in names and n in priv and hasattr(lib, n)


def test_no_cpu_fallback_create_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from hpe_b200.tf_smpl.batch_smpl import SMPL
    with pytest.raises(_lib.SmplbError) as ei:
        SMPL(synthetic.make_model(num_verts=40, regressor_nnz=4))
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "human-pose-estimation_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle[/.]\w+|#include.*oracle", txt, flags=re.M) \
                   , "%s uses the oracle" % f
                assert "import torch" not in txt, "%s imports torch" % f


def test_tf_adapter_is_import_guarded():
    from hpe_b200 import tf_adapter
    try:
        import tensorflow  # noqa: F401
        pytest.skip("TensorFlow is installed")
    except ImportError:
        pass
    with pytest.raises(ImportError):
        tf_adapter.make_tf_smpl(object())


def test_bad_joint_type_raises():
    from hpe_b200.tf_smpl.batch_smpl import load_model_arrays
    with pytest.raises(ValueError):
        load_model_arrays(synthetic.make_model(num_verts=40, regressor_nnz=4), "coco")


def test_model_relayout_matches_reference_loader(small_model):
    """load_model_arrays == what SMPL.__init__ of the reference builds
    (batch_smpl.py:34-81), checked through the oracle's restatement."""
    from hpe_b200.tf_smpl.batch_smpl import load_model_arrays
    from oracle import smpl_numpy as onp
    m = load_model_arrays(small_model, "lsp")
    o = onp.SMPL(small_model, "lsp", dtype=np.float32)
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights", "joint_regressor"):
        assert np.array_equal(m[k], getattr(o, k)), k
    assert np.array_equal(m["parents"], o.parents) and m["parents"][0] == -1
    assert m["joint_regressor"].shape[1] == 14


def test_silhouette_csr_matches_reference_selection():
    """(n,row,col) list -> per-image (x,y) = (col,row) in original order
    (ops.py:123-125)."""
    seg = synthetic.make_silhouettes(4, a_range=(3, 5), b_range=(4, 7))
    pts3 = synthetic.silhouette_points(seg)
    pts, offs = ops.silhouette_csr(pts3, 4)
    assert offs[0] == 0 and offs[-1] == len(pts3) and offs[3] == offs[2]       # image 2 is empty
    for i in range(4):
        rows = pts3[pts3[:, 0] == i]
        want = np.stack([rows[:, 2], rows[:, 1]], axis=1)
        assert np.array_equal(pts[offs[i]:offs[i + 1]], want)
    # unsorted input is regrouped stably
    perm = np.random.default_rng(0).permutation(len(pts3))
    pts_u, offs_u = ops.silhouette_csr(pts3[perm], 4)
    assert np.array_equal(offs_u, offs)
    for i in range(4):
        rows = pts3[perm][pts3[perm][:, 0] == i]
        assert np.array_equal(pts_u[offs[i]:offs[i + 1]], np.stack([rows[:, 2], rows[:, 1]], axis=1))


def test_synthetic_pkl_roundtrip(tmp_path):
    m = synthetic.make_model(num_verts=50, regressor_nnz=5)
    p = tmp_path / "model.pkl"
    synthetic.write_pkl(m, str(p))
    import pickle
    dd = pickle.load(open(str(p), "rb"))
    assert set(dd) == {"v_template", "shapedirs", "posedirs", "J_regressor", "cocoplus_regressor", "weights",
                       "kintree_table"}
    assert dd["kintree_table"][0][0] == 4294967295 and dd["kintree_table"].dtype == np.uint32
