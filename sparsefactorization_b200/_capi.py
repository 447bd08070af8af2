"""ctypes binding of libpsf_b200.so (include/psf_b200.h).  No CPU fallback: if the library is
missing and cannot be built, importing the operators fails loudly."""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

PSF_OK = 0
PSF_F32, PSF_F64 = 0, 1
PATH_AUTO, PATH_GENERIC, PATH_FUSED, PATH_FUSED_GEN1 = 0, 1, 2, 3
MAX_TAPS = 32


class WStrides(C.Structure):
    _fields_ = [("factor", C.c_int64), ("sample", C.c_int64), ("row", C.c_int64), ("tap", C.c_int64)]


class ChainDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("n", C.c_int64), ("C", C.c_int64),
        ("n_W", C.c_int), ("L", C.c_int), ("residual", C.c_int), ("dtype", C.c_int),
        ("ws", WStrides),
        ("offsets", C.POINTER(C.c_int64)),
        ("path", C.c_int),
        ("interleave", C.c_int),
    ]


class GenDesc(C.Structure):
    _fields_ = [("rows", C.c_int64), ("n", C.c_int64), ("n_W", C.c_int), ("E", C.c_int), ("H", C.c_int), ("L", C.c_int),
                ("stride_factor", C.c_int64), ("stride_tap", C.c_int64), ("interleave", C.c_int), ("C", C.c_int)]


class MlpPtrs(C.Structure):       # psf_mlp_params / psf_mlp_grads (same shape)
    _fields_ = [("W1", C.c_void_p), ("b1", C.c_void_p), ("W2", C.c_void_p), ("b2", C.c_void_p)]


_lock = threading.Lock()
_lib = None

_SIGNATURES = {
    "psf_version": (C.c_int, []),
    "psf_error_string": (C.c_char_p, [C.c_int]),
    "psf_chord_offsets": (C.c_int, [C.c_int64, C.c_int, C.POINTER(C.c_int64)]),
    "psf_chord_indices_host": (C.c_int, [C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "psf_chord_indices_device": (C.c_int, [C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "psf_chord_mask_host": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "psf_factor_fwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int64,
                                 C.c_int64, C.c_int, C.c_int64, C.c_void_p]),
    "psf_factor_bwd": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_int64), C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_void_p]),
    "psf_chain_workspace_bytes": (C.c_size_t, [C.POINTER(ChainDesc), C.c_int]),
    "psf_chain_fwd": (C.c_int, [C.POINTER(ChainDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "psf_chain_bwd": (C.c_int, [C.POINTER(ChainDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_size_t, C.POINTER(C.c_int), C.c_void_p]),
    "psf_w_relayout": (C.c_int, [C.c_int, C.c_void_p, WStrides, C.c_void_p, WStrides, C.c_int, C.c_int64, C.c_int64,
                                 C.c_int, C.c_void_p]),
    "psf_w_relayout_ex": (C.c_int, [C.c_int, C.c_void_p, WStrides, C.c_int, C.c_void_p, WStrides, C.c_int, C.c_int,
                                    C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "psf_chain_preferred_interleave": (C.c_int, [C.POINTER(ChainDesc)]),
    "psf_chain_fused_generation": (C.c_int, [C.POINTER(ChainDesc)]),
    "psf_generators_workspace_bytes": (C.c_size_t, [C.c_int]),
    "psf_generators_fwd": (C.c_int, [C.POINTER(GenDesc), C.c_void_p, MlpPtrs, MlpPtrs, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "psf_generators_bwd": (C.c_int, [C.POINTER(GenDesc), C.c_void_p, MlpPtrs, MlpPtrs, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, MlpPtrs, MlpPtrs, C.c_void_p, C.c_size_t, C.c_void_p]),
    "psf_chain_fused_plan": (C.c_int, [C.POINTER(ChainDesc), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "psf_init_linear_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "psf_init_linear_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "psf_init_linear_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "psf_head_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "psf_head_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "psf_head_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                               C.c_void_p, C.c_size_t, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def _build_failed_without_nvcc(exc) -> bool:
    return "nvcc not found" in str(exc)


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (rebuilding first if the sources changed since the library was built) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("PSF_B200_LIB")   # override: A/B runs of differently tuned builds
        if not path:
            try:
                path = _build.build()            # no-op unless the sources changed since the library was built
            except Exception as exc:             # no nvcc on the box
                if os.path.exists(_build.LIB_PATH) and _build_failed_without_nvcc(exc):
                    path = _build.LIB_PATH       # a prebuilt library and no compiler: use it
                else:
                    raise ImportError(
                        "libpsf_b200.so is missing or stale and could not be built; run "
                        "`python -m sparsefactorization_b200.build` where nvcc is available. "
                        "There is no CPU fallback for the PSF chain.") from exc
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


class PsfError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status != PSF_OK:
        msg = load().psf_error_string(status).decode()
        raise PsfError(f"{what} failed with status {status}: {msg}")
