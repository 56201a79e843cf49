"""B200-native episodic-memory retrieval path for Aura (auralmn/aura-snn-rag).

Drop-in for the reference's `src/core/hippocampal.py` memory bank + centroid index; all
arithmetic runs in libaura_hippo.so (hand-written sm_100a CUDA behind a C ABI).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
