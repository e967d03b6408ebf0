"""mojo_simdjson_b200 -- B200 (sm_100a) stage-1 structural indexer behind mojo-simdjson's stage-1 API.

Host-side mirror of the reference interface for this path.  The compute is libsimdjson_b200.so
(hand-written CUDA, C ABI in include/simdjson_b200.h); importing this package never falls back to a CPU
implementation -- if the library or a CUDA device is missing, calls raise.
"""
from . import errors  # noqa: F401

__all__ = ["errors"]
