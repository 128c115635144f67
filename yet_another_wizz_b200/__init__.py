"""
yet_another_wizz_b200 -- B200-native pair-counting engine behind the public API
of yet_another_wizz (`crosscorrelate` / `autocorrelate`).

Only the hot path lives here: CUDA kernels + C ABI under
`csrc/`, and the thin Python host that mirrors the reference's interface for
that path.  There is no CPU fallback.
"""

from ._lib import YawbError
from .engine import DeviceCatalog, Engine

__all__ = ["DeviceCatalog", "Engine", "YawbError"]
__version__ = "0.1.0"
