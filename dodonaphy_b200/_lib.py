"""ctypes binding of the C ABI in include/dodonaphy_b200.h.

There is no CPU fallback: if the shared library is missing and cannot be built, or CUDA is not
available when a kernel entry point is called, the call raises.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# DODO_LIB: load another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("DODO_LIB") or os.path.join(_PKG, "_C", "libdodonaphy_b200.so")

DODO_OK, DODO_EINVAL, DODO_ECUDA, DODO_ENOTIMPL, DODO_EINEXACT = 0, -1, -2, -3, -4
MODEL_JC69, MODEL_EIGEN = 0, 1
PRUNE_GRAD, PRUNE_MIX, PRUNE_SCALE, PRUNE_GMATS, PRUNE_GMODEL = 1, 2, 4, 8, 16
PDM_TORCH, PDM_NUMPY = 0, 1
LIFT_UP, LIFT_WRAP = 0, 1
DIST_POINCARE, DIST_LORENTZ, DIST_SHEET_UP = 0, 1, 2
PRIOR_TORCH, PRIOR_NUMPY = 0, 1


class Model(C.Structure):
    _fields_ = [("U", C.c_double * 16), ("Uinv", C.c_double * 16), ("lam", C.c_double * 4),
                ("Q", C.c_double * 16), ("pi", C.c_double * 4), ("kind", C.c_int32), ("_pad", C.c_int32)]


class Prior(C.Structure):
    _fields_ = [("aT", C.c_double), ("bT", C.c_double), ("a", C.c_double), ("c", C.c_double), ("variant", C.c_int32),
                ("_pad", C.c_int32), ("d_ln_prior", C.c_void_p), ("d_grad_blens", C.c_void_p)]


class PrunePlan(C.Structure):
    _fields_ = [("B", C.c_int32), ("S", C.c_int32), ("L", C.c_int32), ("K", C.c_int32),
                ("flags", C.c_uint32), ("threads", C.c_int32), ("grid", C.c_int32),
                ("tile_patterns", C.c_int32), ("n_tiles", C.c_int32), ("n_items", C.c_int32),
                ("cherry_slots", C.c_int32), ("ws_bytes", C.c_uint64),
                ("off_ops", C.c_uint64), ("off_sched", C.c_uint64), ("off_pmat", C.c_uint64),
                ("off_tiptab", C.c_uint64), ("off_scratch", C.c_uint64), ("off_ll_part", C.c_uint64),
                ("off_g_part", C.c_uint64), ("off_gm_part", C.c_uint64)]


_SIGNATURES = {
    "dodo_last_error": (C.c_char_p, []),
    "dodo_version": (C.c_int, []),
    "dodo_launch_count": (C.c_uint64, []),
    "dodo_set_timing_events": (None, [C.c_void_p, C.c_void_p]),
    "dodo_prune_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                  C.POINTER(PrunePlan)]),
    "dodo_prune": (C.c_int, [C.POINTER(PrunePlan), C.POINTER(Model), C.POINTER(C.c_double), C.c_void_p,
                             C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dodo_prune_model": (C.c_int, [C.POINTER(PrunePlan), C.POINTER(Model), C.POINTER(C.c_double), C.c_void_p,
                                   C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dodo_pdm_workspace_bytes": (C.c_uint64, [C.c_int, C.c_int, C.c_int]),
    "dodo_pdm_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                               C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dodo_pdm_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                               C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dodo_model_chain": (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "dodo_blens_jacobian_workspace_bytes": (C.c_uint64, [C.c_int, C.c_int, C.c_int]),
    "dodo_blens_jacobian": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dodo_gram_logdet": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dodo_alignment_workspace_bytes": (C.c_uint64, [C.c_int, C.c_int64]),
    "dodo_alignment_sort": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dodo_alignment_gather": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "dodo_hypdist_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dodo_hypdist_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dodo_nj_workspace_bytes": (C.c_uint64, [C.c_int, C.c_int]),
    "dodo_prior_gamma_dir": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(Prior), C.c_void_p]),
    "dodo_nj_hard": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.POINTER(Prior), C.c_void_p]),
    "dodo_nj_soft": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.POINTER(Prior), C.c_void_p]),
    "dodo_nj_guard": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dodo_soft_argmin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "dodo_nj_soft_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "dodo_geo_workspace_bytes": (C.c_uint64, [C.c_int, C.c_int, C.c_int]),
    "dodo_geo_soft": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 8),
    "dodo_geo_soft_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dodo_geo_hard": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load(build_if_missing=True):
    """Load (building first if necessary) the CUDA library.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("DODO_LIB", LIB_PATH)  # an alternative build of the same library (A/B timing runs)
    if path != LIB_PATH:
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m dodonaphy_b200.build`")
        from . import build as _build

        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DodoError(RuntimeError):
    pass


def check(code):
    """Turn a status code into the exception type the reference raises."""
    if code == DODO_OK:
        return
    msg = load().dodo_last_error().decode("utf-8", "replace")
    if code == DODO_EINVAL:
        raise ValueError(msg)
    if code == DODO_ENOTIMPL:
        raise NotImplementedError(msg)
    raise DodoError(f"[{code}] {msg}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("dodonaphy_b200 has no CPU path: a CUDA device (B200, sm_100a) is required")


def launch_count():
    return int(load().dodo_launch_count())
