"""CPU-side checks of the drop-in boundary: the C-ABI library loads here (no GPU, no compute calls)
and exports every symbol include/aura_hippo.h declares; the ctypes table matches the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "aura_hippo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aura_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from aura_snn_rag_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from aura_snn_rag_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/aura_hippo.h but not exported"


def test_ctypes_table_matches_header():
    from aura_snn_rag_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    lib = _lib.load()
    assert lib.aura_version() == 300
    assert lib.aura_last_error_string() is not None


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call, with a message."""
    from aura_snn_rag_b200 import _lib
    lib = _lib.load()
    rc = lib.aura_scan_topk(None, 7, 10, 4, None, 1, None, None, 5, 0, None, None, None, 0, None)
    assert rc == -1 and b"dtype" in lib.aura_last_error_string()
    rc = lib.aura_scan_topk(None, 0, 10, 4, None, 1, None, None, 1000, 0, None, None, None, 0, None)
    assert rc == -1 and b"k=1000" in lib.aura_last_error_string()
    rc = lib.aura_topk_merge(None, None, 2, 0, 1, 1, None, None, None)
    assert rc == -1
    with pytest.raises(_lib.AuraLibraryError):
        _lib.check(rc, "aura_topk_merge")
    assert lib.aura_scan_topk_workspace_bytes(1000, 64, 4, 10) > 0
    # list-major copies: the bank's element type, or a bf16 shadow of an fp32 bank (with its rounding-error scalar)
    F32, BF16 = _lib.AURA_F32, _lib.AURA_BF16
    rc = lib.aura_ivf_pack_lists(None, BF16, 64, None, 10, None, F32, None, None)
    assert rc == -1 and b"shadow" in lib.aura_last_error_string()
    one = ctypes.c_void_p(256)                        # any non-null, 16-byte aligned address: validation only, never dereferenced
    rc = lib.aura_ivf_search_batch(one, F32, 1000, 64, one, 64, one, 16, 4, one, one, one, BF16, None, None, None, 10, 0, 0,
                                   0.5, one, one, one, one, 1 << 40, None)
    assert rc == -1 and b"rounding-error" in lib.aura_last_error_string()
    rc = lib.aura_ivf_search_batch(one, BF16, 1000, 64, one, 64, one, 16, 4, one, one, one, F32, None, None, None, 10, 0, 0,
                                   0.5, one, one, one, one, 1 << 40, None)
    assert rc == -1 and b"list-major copy" in lib.aura_last_error_string()
    # the watchdog trace words exist (all zero: nothing has trapped)
    out = (ctypes.c_uint32 * 4)(1, 1, 1, 1)
    assert lib.aura_debug_last_trap(out) == 0 and list(out) == [0, 0, 0, 0]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from aura_snn_rag_b200 import _lib
    from aura_snn_rag_b200.hippocampal import HippocampalFormation
    with pytest.raises(_lib.AuraLibraryError):
        HippocampalFormation(max_memories=8, feature_dim=4)
    from aura_snn_rag_b200 import ops
    with pytest.raises(_lib.AuraLibraryError):
        ops.row_inv_norms(torch.zeros(4, 4))
