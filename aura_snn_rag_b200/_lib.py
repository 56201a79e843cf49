"""ctypes binding of libaura_hippo.so (C ABI: include/aura_hippo.h).

There is exactly one implementation of the hot path: the CUDA library.  If it is missing or
a call fails, this module raises - there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libaura_hippo.so")

AURA_F32, AURA_BF16 = 0, 1
AURA_MAX_K = 128
AURA_MAX_NPROBE = 128
AURA_IVF_EMPTY_OK = 1
AURA_IVF_MEASURED_EPS = 2

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); kept in the order of include/aura_hippo.h
SIGNATURES = {
    "aura_version": (_i, []),
    "aura_last_error_string": (C.c_char_p, []),
    "aura_kernel_launches": (C.c_uint64, []),
    "aura_debug_last_trap": (_i, [_p]),
    "aura_row_inv_norms": (_i, [_p, _i, _i64, _i, _p, _p]),
    "aura_row_terms": (_i, [_p, _p, _i, _p, _f, _p, _i64, _p, _p, _p]),
    "aura_decay_strength": (_i, [_p, _i64, _f, _p]),
    "aura_scan_topk_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "aura_scan_topk": (_i, [_p, _i, _i64, _i, _p, _i, _p, _p, _i, _i64, _p, _p, _p, _sz, _p]),
    "aura_bank_write": (_i, [_p, _i, _i, _i64, _i, _p, _p, _i, _p, _p, _f, _p, _p]),
    "aura_kmeans_seed": (_i, [_p, _i, _i, _p, _i, _p, _p]),
    "aura_kmeans_assign_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "aura_kmeans_assign": (_i, [_p, _i, _i64, _i, _p, _i, _p, _p, _p, _i, _p, _p, _sz, _p]),
    "aura_ivf_build_lists_workspace_bytes": (_sz, [_i]),
    "aura_ivf_build_lists": (_i, [_p, _i64, _i, _p, _p, _p, _sz, _p]),
    "aura_kmeans_list_sums": (_i, [_p, _i, _i, _p, _p, _i, _p, _p, _p]),
    "aura_kmeans_finalize": (_i, [_p, _p, _i, _i, _p, _p]),
    "aura_ivf_list_counts": (_i, [_p, _i, _p, _p]),
    "aura_online_assign_workspace_bytes": (_sz, []),
    "aura_online_assign": (_i, [_p, _i, _i, _i64, _i, _p, _i, _p, _p, _p, _i, _p, _sz, _p]),
    "aura_ivf_coarse_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "aura_ivf_coarse": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _sz, _p]),
    "aura_ivf_search_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "aura_ivf_search": (_i, [_p, _i, _i64, _i, _p, _i, _p, _i, _i, _p, _p, _p, _p, _i, _i64, _i, _p, _p, _p, _p, _sz, _p]),
    "aura_ivf_search_batch_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "aura_ivf_search_batch": (_i, [_p, _i, _i64, _i, _p, _i, _p, _i, _i, _p, _p, _p, _i, _p, _p, _p, _i, _i64, _i, _f, _p, _p, _p, _p, _sz, _p]),
    "aura_ivf_pack_lists": (_i, [_p, _i, _i, _p, _i64, _p, _i, _p, _p]),
    "aura_ivf_search_batch_items": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "aura_batch_topk_workspace_bytes": (_sz, [_i64, _i, _i, _i, _i]),
    "aura_batch_topk": (_i, [_p, _i, _i64, _i, _p, _i, _p, _p, _i, _i64, _f, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "aura_rows_to_bf16": (_i, [_p, _i64, _i, _p, _p, _p]),
    "aura_allpairs_topk_workspace_bytes": (_sz, [_i64, _i64, _i, _i, _i]),
    "aura_allpairs_topk": (_i, [_p, _i, _i64, _i, _i64, _i64, _p, _i, _p, _p, _p, _sz, _p]),
    "aura_topk_merge": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "aura_pack_topk": (_i, [_p, _p, _p, _i, _i, _p, _i64, _p, _p]),
    "aura_topk_merge_packed": (_i, [_p, _i, _i, _i, _p, _p, _p, _p]),
    "aura_peer_gather_buffer_bytes": (_sz, [_i, _i, _i]),
    "aura_pack_scatter": (_i, [_p, _p, _p, _i, _i, _p, _i64, _p, _i, _i, _p, _p]),
    "aura_merge_gathered": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "aura_gather_rows": (_i, [_p, _i, _i, _p, _i64, _p, _p]),
    "aura_gather_context": (_i, [_p, _i, _i, _p, _p, _i, _i, _p, _p, _p]),
}


class AuraLibraryError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AuraLibraryError(
            f"{LIB_PATH} not found: build it with `python -m aura_snn_rag_b200.build` "
            "(there is no CPU fallback for the retrieval path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().aura_last_error_string().decode("utf-8", "replace")
        raise AuraLibraryError(f"{what} failed with status {status}: {msg}")
