"""ctypes loader for libb200search.so.  Fails loudly when the library is missing: the
product has no Python / CPU implementation of any kernel."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200_LIB_PATH: an A/B build of the same library (tools/build_variant.sh); never a different implementation
LIB_PATH = os.environ.get("B200_LIB_PATH") or os.path.join(_HERE, "libb200search.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200_search.h")

_lib = None


class LibraryMissing(RuntimeError):
    pass


def declared_symbols():
    """Every function the C header declares (used by the CPU-side export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C myscaledb_b200/csrc).  There is no fallback implementation.")
        L = C.CDLL(LIB_PATH)
        L.b200_last_error.restype = C.c_char_p
        L.b200_version.restype = C.c_char_p
        L.b200_launch_count.restype = C.c_int64
        _lib = L
    return _lib
