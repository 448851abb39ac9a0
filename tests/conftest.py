import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import hpe_b200  # noqa: E402,F401  (alias for human-pose-estimation_b200/)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def small_model_from_golden(g):
    """Rebuild the reference-pickle-layout model dict stored in smpl_small.npz."""
    import scipy.sparse as sp
    return {
        "v_template": g["m_v_template"].astype(np.float64),
        "shapedirs": g["m_shapedirs"].astype(np.float64),
        "posedirs": g["m_posedirs"].astype(np.float64),
        "weights": g["m_weights"].astype(np.float64),
        "J_regressor": sp.csc_matrix(g["m_J_regressor"].astype(np.float64)),
        "cocoplus_regressor": sp.csc_matrix(g["m_cocoplus_regressor"].astype(np.float64)),
        "kintree_table": g["m_kintree_table"],
    }


@pytest.fixture(scope="session")
def golden_small():
    return load_golden("smpl_small.npz")


@pytest.fixture(scope="session")
def golden_small_lsp():
    return load_golden("smpl_small_lsp.npz")


@pytest.fixture(scope="session")
def golden_full():
    return load_golden("smpl_full.npz")


@pytest.fixture(scope="session")
def small_model(golden_small):
    return small_model_from_golden(golden_small)


@pytest.fixture(scope="session")
def full_model():
    from hpe_b200 import synthetic
    return synthetic.make_model(seed=0)


def rel_err(a, ref):
    """max |a - ref| / max |ref| -- the scale-relative error of SURVEY.md §8c."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(a - ref)) / max(float(np.max(np.abs(ref))), 1e-30))


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """libsmplb.so is built in-tree (nvcc cross-compiles without a GPU)."""
    from hpe_b200._lib import LIB_PATH
    if not os.path.isfile(LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return LIB_PATH
