"""B200-native episodic-memory retrieval path for Aura (auralmn/aura-snn-rag).

Drop-in for the reference's `src/core/hippocampal.py` memory bank + centroid index; all
arithmetic runs in libaura_hippo.so (hand-written sm_100a CUDA behind a C ABI,
include/aura_hippo.h).  Importing the package does not need a GPU; constructing
`HippocampalFormation` or calling anything in `ops` does, and fails loudly without one.
"""
from . import _lib  # noqa: F401
from .idtable import IdTable  # noqa: F401

__all__ = ["_lib", "IdTable", "HippocampalFormation", "EpisodicMemory"]


def __getattr__(name):
    if name in ("HippocampalFormation", "EpisodicMemory"):
        from . import hippocampal
        return getattr(hippocampal, name)
    raise AttributeError(name)
