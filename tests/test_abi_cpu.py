"""CPU tests of the boundary: the C-ABI library loads and exports every symbol that
include/maai_ntxent.h declares (no compute calls), and the host mirror validates arguments and
refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

import maai_b200
from maai_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "maai_ntxent.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(maai_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 7
    for n in names:
        assert hasattr(lib, n), f"symbol {n} declared in include/maai_ntxent.h is not exported"
    # and the binding covers exactly the declared set
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_pure_host_helpers():
    lib = _lib.load()
    assert lib.maai_abi_version() == _lib.ABI_VERSION
    assert [lib.maai_padded_dim(d) for d in (1, 20, 64, 65, 128, 129, 256)] == [64, 64, 64, 128, 128, 256, 256]
    assert lib.maai_padded_dim(0) == _lib.E_SHAPE and lib.maai_padded_dim(257) == _lib.E_SHAPE
    assert lib.maai_ntxent_r_len(100, 3) == 640 and lib.maai_ntxent_r_len(64, 1) == 128


def test_c_abi_argument_errors_without_gpu():
    lib = _lib.load()
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.maai_ntxent_normalize(None, None, 4, 8, 0, None, None, None, None, 0, None) == _lib.E_ARG
    assert b"null" in lib.maai_last_error()
    assert lib.maai_ntxent_fwd(None, 4, 1, 0, 64, 1.0, None, None, None, None, 0, None, None) == _lib.E_ARG
    # the later entry points validate the same way (ABI v4 / v5)
    assert lib.maai_ntxent_bwd_tiles(None, None, None, 4, 1, 0, 64, 1.0, 3, None, None) == _lib.E_ARG
    assert lib.maai_ntxent_bwd_keyside(None, None, 4, 2, 0, 64, 1.0, None, None) == _lib.E_ARG
    assert lib.maai_ntxent_bwd_dh(None, None, None, None, None, None, 0, None, None, 4, 8, 64, 1.0, 1, 3,
                                  None, None, None) == _lib.E_ARG
    assert lib.maai_ntxent_fwd_sym_tiles(None, 4, 2, 0, 64, 1.0, None, None, 0, None, None) == _lib.E_ARG
    assert lib.maai_ntxent_fwd_sym_finalize(None, None, 4, 2, 0, 1.0, None, None, None, None, None, None, None) == _lib.E_ARG
    import ctypes
    buf = (ctypes.c_float * 64)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    # bad rank / world / width with non-null pointers: still no CUDA call
    assert lib.maai_ntxent_fwd_sym_tiles(ptr, 4, 2, 2, 64, 1.0, ptr, ptr, 0, None, None) == _lib.E_ARG
    assert lib.maai_ntxent_fwd_sym_tiles(ptr, 4, 17, 0, 64, 1.0, ptr, ptr, 0, None, None) == _lib.E_SHAPE
    assert lib.maai_ntxent_bwd_keyside(ptr, ptr, 4, 2, 0, 96, 1.0, ptr, None) == _lib.E_SHAPE
    # ABI v6: step workspace, flags, zero-fill arguments
    assert lib.maai_ntxent_workspace_bytes(100, 128, 0) == 256 * 4
    assert lib.maai_ntxent_workspace_bytes(100, 128, 1) == (256 + 200 * 128) * 4
    assert lib.maai_ntxent_workspace_bytes(100, 96, 1) == 0 and lib.maai_ntxent_workspace_bytes(0, 128, 1) == 0
    assert lib.maai_ntxent_fwd(ptr, 4, 1, 0, 64, 1.0, ptr, ptr, None, ptr, 2, None, None) == _lib.E_ARG  # unknown flag
    assert b"flag" in lib.maai_last_error()
    assert lib.maai_ntxent_normalize(ptr, ptr, 4, 8, 0, ptr, ptr, ptr, ptr, 6, None) == _lib.E_ARG  # zero_bytes % 4
    assert lib.maai_ntxent_normalize(ptr, ptr, 4, 8, 0, ptr, ptr, ptr, None, 16, None) == _lib.E_ARG
    assert lib.maai_ntxent_fwd_is_symmetric(4096, 1, 128) == 1 and lib.maai_ntxent_fwd_is_symmetric(4096, 2, 128) == 0
    assert lib.maai_ntxent_fwd_is_symmetric(2048, 1, 256) == 0 and lib.maai_ntxent_fwd_is_symmetric(8192, 1, 256) == 1
    # ABI v7: maai_peer_sync is validated on the host (null members, seq 0) before any CUDA call
    bad = _lib.PeerSync(None, None, None, 1, 0)
    assert lib.maai_ntxent_fwd(ptr, 4, 2, 0, 64, 1.0, ptr, ptr, None, ptr, 0, ctypes.byref(bad), None) == _lib.E_ARG
    assert b"maai_peer_sync" in lib.maai_last_error()
    bad = _lib.PeerSync(ptr.value, ptr.value, ptr.value, 0, 0)
    assert lib.maai_ntxent_bwd(ptr, ptr, ptr, 1, ptr, ptr, ptr, ptr, 0, ptr, ptr, 4, 2, 0, 8, 64, 1.0, 3, ptr, ptr, ptr, 0,
                               ctypes.byref(bad), None) == _lib.E_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.E_SHAPE, "x")
    with pytest.raises(_lib.MaaiError):
        _lib.check(_lib.E_CUDA, "x")


def test_host_mirror_signature_matches_reference():
    import inspect
    sig = inspect.signature(maai_b200.contrastive_loss)
    names = list(sig.parameters)
    # Objective.py:17-22
    assert names[:7] == ["hidden1", "hidden2", "hidden_norm", "temperature", "local_rank", "world_size", "device"]
    p = sig.parameters
    assert p["hidden_norm"].default is True and p["temperature"].default == 1.0
    assert p["local_rank"].default == 0 and p["world_size"].default == 1 and p["device"].default == "cpu"


def test_no_cpu_fallback_and_argument_validation():
    a, b = torch.randn(4, 8), torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        maai_b200.contrastive_loss(a, b)
    with pytest.raises(AssertionError):  # Objective.py:45
        maai_b200.contrastive_loss(a, torch.randn(5, 8))
    with pytest.raises(NotImplementedError):
        maai_b200.contrastive_loss(a, b, hidden_norm=False)


def test_deterministic_mode_is_honoured():
    """fp32 atomics across CTAs: like torch's own non-deterministic CUDA ops, raise under
    torch.use_deterministic_algorithms(True) and warn once under warn_only."""
    a, b = torch.randn(4, 8), torch.randn(4, 8)
    try:
        torch.use_deterministic_algorithms(True)
        with pytest.raises(RuntimeError, match="deterministic implementation"):
            maai_b200.contrastive_loss(a, b)
        torch.use_deterministic_algorithms(True, warn_only=True)
        with pytest.warns(UserWarning, match="deterministic implementation"):
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                maai_b200.contrastive_loss(a, b)
    finally:
        torch.use_deterministic_algorithms(False)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "multimodal-active-ai_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU or PyTorch fallback", ""), f"{f} mentions oracle"


def test_integration_doc_stub_matches_the_abi():
    """INTEGRATION.md shows the ctypes stub a maintainer would write: its argument lists must have the length of
    the real signatures (the ABI changed three times this round; the doc must not drift)."""
    import re
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    found = dict(re.findall(r"lib\.(maai_\w+)\.argtypes\s*=\s*\[([^\]]*)\]", text))
    assert {"maai_ntxent_normalize", "maai_ntxent_fwd", "maai_ntxent_bwd"} <= set(found)
    for name, args in found.items():
        n_doc = len([a for a in args.split(",") if a.strip()])
        assert n_doc == len(_lib.SIGNATURES[name][1]), (name, n_doc, len(_lib.SIGNATURES[name][1]))
    assert f"ABI v{_lib.ABI_VERSION}" in text
