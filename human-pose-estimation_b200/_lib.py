"""ctypes binding of libsmplb.so (include/smplb.h).  The library is the product:
if it is missing this module raises -- there is no Python/CPU fallback."""
import ctypes as C
import os

HOST, DEVICE = 0, 1
GP_FLOATS = 428
NUM_JOINTS = 24

_HERE = os.path.dirname(os.path.abspath(__file__))
# (SMPLB_LIB: an instrumented build of the same library, e.g. -DFB_TIMING, for the tools/)
LIB_PATH = os.environ.get("SMPLB_LIB") or os.path.join(_HERE, "libsmplb.so")


class SmplbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsmplb error %d: %s" % (code, msg))
        self.code = code


class Model(C.Structure):
    _fields_ = [("num_verts", C.c_int32), ("num_betas", C.c_int32), ("num_keypoints", C.c_int32),
                ("reserved", C.c_int32), ("v_template", C.c_void_p), ("shapedirs", C.c_void_p),
                ("posedirs", C.c_void_p), ("J_regressor", C.c_void_p), ("weights", C.c_void_p),
                ("joint_regressor", C.c_void_p), ("parents", C.c_void_p)]


_P, _I, _F, _L, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t

# name -> argtypes, in the order of include/smplb.h
SIGNATURES = {
    "smplb_create": [C.POINTER(_P), C.POINTER(Model), _I, _I],
    "smplb_destroy": [_P],
    "smplb_malloc": [_P, C.POINTER(_P), _SZ],
    "smplb_free": [_P, _P],
    "smplb_host_alloc": [C.POINTER(_P), _SZ],
    "smplb_host_free": [_P],
    "smplb_memcpy_h2d": [_P, _P, _P, _SZ],
    "smplb_memcpy_d2h": [_P, _P, _P, _SZ],
    "smplb_memset": [_P, _P, _I, _SZ],
    "smplb_sync": [_P],
    "smplb_order_after": [_P, _P],
    "smplb_flush_l2": [_P, _SZ],
    "smplb_timer_start": [_P, _I],
    "smplb_timer_stop": [_P, _I],
    "smplb_timer_elapsed_ms": [_P, _I, C.POINTER(_F)],
    "smplb_launch_count": [_P, C.POINTER(_L)],
    "smplb_profile_enable": [_P, _I],
    "smplb_profile_read": [_P, C.c_char_p, _SZ],
    "smplb_smpl_forward": [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I],
    "smplb_smpl_backward": [_P, _I, _P, _P, _P, _P, _P, _I],
    "smplb_last_verts": [_P, C.POINTER(_P)],
    "smplb_rodrigues": [_P, _I, _P, _P, _I],
    "smplb_global_rigid": [_P, _I, _P, _P, _P, _P, _I],
    "smplb_skew": [_P, _I, _P, _P, _I],
    "smplb_lrotmin": [_P, _I, _P, _P, _I],
    "smplb_orth_proj": [_P, _I, _I, _P, _P, _P, _I],
    "smplb_reproject_vertices": [_P, _I, _I, _P, _P, _F, _F, _P, _I],
    "smplb_proj_backward": [_P, _I, _I, _P, _P, _P, _I, _F, _F, _P, _P, _I],
    "smplb_kp_loss": [_P, _I, _I, _P, _P, _P, _P, _P, _I],
    "smplb_mesh_reproj_loss": [_P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _I],
    "smplb_gradient_penalty": [_P, _I, _P, _P, _P, _P, _P, _P, _I],
    "smplb_gradient_penalty_from_sums": [_P, _L, _P, _P, _I],
    "smplb_gradient_penalty_backward": [_P, _I, _L, _P, _P, _P, _P, _P, _I],
    "smplb_silhouette_csr": [_P, _I, _I, _I, _P, _P, _I, _P, _I],
    "smplb_kcs": [_P, _I, _I, _P, _P, _P, _I],
    "smplb_kcs_backward": [_P, _I, _I, _P, _P, _P, _P, _I],
    "smplb_interpolate": [_P, _I, _I, _P, _P, _P, _P, _I],
    "smplb_critic_inputs": [_P, _I, _I] + [_P] * 14 + [_I],
    "smplb_critic_gradient_penalty": [_P, _I, _I, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I],
    "smplb_step": [_P, _I, _P, _P, _P, _P, _P, _P, _I, _F, _F, _F, _L, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I],
    "smplb_step_seg": [_P, _I, _P, _P, _P, _P, _P, _I, _I, _F, _F, _F, _L, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I],
    "smplb_comm_p2p_export": [_P, _P],
    "smplb_comm_p2p_attach": [_P, _I, _I, _P],
    "smplb_comm_p2p_attach_local": [_P, _I, _I, C.POINTER(_P)],
    "smplb_comm_status": [_P, C.POINTER(_I)],
    "smplb_comm_unique_id": [_P],
    "smplb_comm_init": [_P, _I, _I, _P],
    "smplb_comm_allreduce_sum": [_P, _P, _I],
    "smplb_comm_destroy": [_P],
}

# private hooks (csrc/smplb_debug.h), not part of include/smplb.h
PRIVATE_SIGNATURES = {
    "smplb_debug_set": [_P, C.c_char_p, _I],
    "smplb_debug_p2p_inject": [_P, _I, C.c_uint, _I, _F, _F, C.c_longlong],
}
STEP_KEEP_VERTS = 1

_lib = None


def _point_at_bundled_nccl():
    """smplb_comm_init dlopens NCCL.  If this interpreter has a PyTorch with its own libnccl.so.2 (the nvidia-nccl wheel),
    name THAT file in SMPLB_NCCL_LIB: the first libnccl.so.2 mapped into a process serves every later request for the
    soname, and a torch imported after a comm_init that picked the (older) system library fails to resolve its
    symbols.  No torch import here; a process without the wheel keeps the system library."""
    if os.environ.get("SMPLB_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.isfile(cand):
                os.environ["SMPLB_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def lib():
    """The loaded library; raises if libsmplb.so has not been built
    (`python -c "import __graft_entry__ as g; g.build()"` or `make -C csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError("libsmplb.so not found at %s: build it with __graft_entry__.build(); "
                              "there is no CPU fallback" % LIB_PATH)
        _point_at_bundled_nccl()
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, args in list(SIGNATURES.items()) + list(PRIVATE_SIGNATURES.items()):
            f = getattr(l, name)
            f.argtypes = args
            f.restype = C.c_int
        l.smplb_last_error.restype = C.c_char_p
        l.smplb_last_error.argtypes = []
        l.smplb_version.restype = C.c_int
        l.smplb_version.argtypes = []
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise SmplbError(rc, lib().smplb_last_error().decode("utf-8", "replace"))
