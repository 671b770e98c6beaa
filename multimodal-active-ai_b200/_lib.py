"""ctypes binding of include/maai_ntxent.h.  There is no fallback: if the CUDA library is missing
or a call fails, the caller gets an exception."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MAAI_DEBUG_LIB selects an A/B build of the same library (tools/ab_variants.py); never a fallback
LIB_PATH = os.environ.get("MAAI_DEBUG_LIB") or os.path.join(HERE, "libmaai_ntxent.so")

ABI_VERSION = 8
F_PREZEROED = 1
WS_CTL_WORDS = 32
OK, E_ARG, E_SHAPE, E_CUDA = 0, -1, -2, -3
DT_F32, DT_BF16, DT_F16 = 0, 1, 2

_c_int, _c_float, _c_void_p, _c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
FLAG_WORDS = 96


class PeerSync(ctypes.Structure):
    """maai_peer_sync of include/maai_ntxent.h (host struct handed to the calls by pointer)."""
    _fields_ = [("peer_flag_bases", _c_void_p), ("local_flags", _c_void_p), ("counter", _c_void_p),
                ("seq", ctypes.c_uint), ("timeout_s", ctypes.c_uint)]


_c_sync_p = ctypes.POINTER(PeerSync)

# name -> (restype, argtypes): every symbol declared in include/maai_ntxent.h
SIGNATURES = {
    "maai_abi_version": (_c_int, []),
    "maai_last_error": (ctypes.c_char_p, []),
    "maai_padded_dim": (_c_int, [_c_int]),
    "maai_ntxent_r_len": (_c_size_t, [_c_int, _c_int]),
    "maai_ntxent_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "maai_ntxent_fwd_is_symmetric": (_c_int, [_c_int, _c_int, _c_int]),
    "maai_ntxent_normalize": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p,
                                       _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "maai_ntxent_fwd": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p,
                                 _c_void_p, _c_void_p, _c_void_p, _c_int, _c_sync_p, _c_void_p]),
    "maai_ntxent_normalize_peer": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                                            _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_sync_p,
                                            _c_void_p]),
    "maai_ntxent_normalize_chain": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                             _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                             _c_size_t, _c_sync_p, _c_void_p]),
    "maai_ntxent_fwd_peer": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p,
                                      _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_sync_p, _c_void_p]),
    "maai_ntxent_fwd_eval": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p,
                                      _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "maai_ntxent_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                 _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                 _c_float, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_sync_p, _c_void_p]),
    "maai_ntxent_fwd_sym_tiles": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p, _c_void_p,
                                           _c_int, _c_sync_p, _c_void_p]),
    "maai_ntxent_fwd_sym_finalize": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_void_p,
                                              _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_sync_p, _c_void_p]),
    "maai_ntxent_fwd_sym_direct": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p, _c_void_p,
                                            _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_sync_p, _c_void_p]),
    "maai_ntxent_bwd_tiles": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float,
                                       _c_int, _c_void_p, _c_void_p]),
    "maai_ntxent_bwd_keyside": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float,
                                         _c_void_p, _c_void_p]),
    "maai_ntxent_bwd_dh": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                                    _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_int, _c_int,
                                    _c_void_p, _c_void_p, _c_void_p]),
    "maai_debug_group_plan": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                       _c_void_p]),
    "maai_debug_tri_locate": (_c_int, [ctypes.c_longlong, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "maai_launch_count": (ctypes.c_ulonglong, []),
}

_lib = None
_ext = None  # False once found missing
EXT_PATH = os.path.join(HERE, "maai_torch_ext.so")


class MaaiError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libmaai_ntxent.so (built by build.py / __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MaaiError(
            f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
            "(run `python multimodal-active-ai_b200/build.py`). There is no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.maai_abi_version() != ABI_VERSION:
        raise MaaiError(f"ABI mismatch: library {lib.maai_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def fast_ext():
    """The C++ autograd binding of the single-rank step (csrc/maai_torch_ext.cpp, built by
    build.py::build_torch_ext), or None when it has not been built.  It enqueues exactly the C-ABI calls the
    Python autograd.Function does -- same kernels, same library instance -- at a fraction of the host cost;
    MAAI_FAST_EXT=0 keeps the Python path (A/B runs).  Never a fallback for the CUDA library itself."""
    global _ext
    if _ext is None:
        _ext = False
        if os.environ.get("MAAI_FAST_EXT", "1") != "0" and os.path.exists(EXT_PATH):
            load()  # libmaai_ntxent.so first: the binding links against it (soname match / rpath $ORIGIN)
            import importlib.util
            import torch  # noqa: F401  (the binding needs libtorch loaded)
            spec = importlib.util.spec_from_file_location("maai_torch_ext", EXT_PATH)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            if mod.abi_version() != ABI_VERSION:
                raise MaaiError(f"maai_torch_ext.so was built against ABI {mod.abi_version()}, binding expects {ABI_VERSION}")
            _ext = mod
    return _ext or None


def check(rc: int, what: str) -> None:
    if rc == OK:
        return
    msg = load().maai_last_error().decode("utf-8", "replace")
    if rc in (E_ARG, E_SHAPE):
        raise ValueError(f"{what}: {msg}")
    raise MaaiError(f"{what}: {msg}")
