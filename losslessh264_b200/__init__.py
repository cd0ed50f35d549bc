"""losslessh264_b200 — B200-native (sm_100a) hot path of losslessh264.

The product is the C-ABI shared library `_lib/liblh264_b200.so` (include/lh264_b200.h); this
package is the thin Python host layer over it (ctypes), used by the tests and bench.py.
"""
from .build import build_library, library_path  # noqa: F401
