"""Deterministic synthetic JSON workloads (BASELINE.json configs 2-4); see gen.c."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import build as _build

SEED_TWITTER = 0x5EED0001
SEED_DOC = 0x5EED0002
SEED_NDJSON = 0x5EED0003
TWITTER_BYTES = 631_515

_lib = None


def _L():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_synth())
        for name in ("sjb200_gen_status_array", "sjb200_gen_ndjson", "sjb200_gen_twitter_pretty"):
            f = getattr(_lib, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    return _lib


def _gen(name: str, size: int, seed: int, out: np.ndarray | None) -> np.ndarray:
    if out is None:
        out = np.empty(size, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= size and out.flags["C_CONTIGUOUS"]
    rc = getattr(_L(), name)(out.ctypes.data, size, seed)
    if rc != 0:
        raise ValueError(f"{name}: size {size} too small")
    return out[:size]


def twitter_like(size: int = TWITTER_BYTES, seed: int = SEED_TWITTER, out=None) -> np.ndarray:
    """Pretty-printed twitter-like document of exactly `size` bytes (config 2)."""
    return _gen("sjb200_gen_twitter_pretty", size, seed, out)


def status_array(size: int, seed: int = SEED_DOC, out=None) -> np.ndarray:
    """One minified array of status objects, exactly `size` bytes (config 3)."""
    return _gen("sjb200_gen_status_array", size, seed, out)


def ndjson(size: int, seed: int = SEED_NDJSON, out=None) -> np.ndarray:
    """One minified status object per line, exactly `size` bytes of whole lines (config 4)."""
    return _gen("sjb200_gen_ndjson", size, seed, out)


def plant_backslash_runs(doc: np.ndarray, every: int = 1 << 20, run: int = 33, chunk: int = 2048) -> int:
    """Adversarial variant of a generated document (config 5): about every `every` bytes, `run` backslashes ending exactly at
    a 2 KiB boundary -- the input for which a 32-byte look-behind cannot decide the escape state entering the next chunk.
    The run replaces ASCII bytes that hold neither a quote nor a backslash (and is followed by a non-quote byte), so the quote
    parity of the document and its UTF-8 validity, hence its stage-1 verdicts, are unchanged.  In place; returns the number of
    runs planted."""
    planted = 0
    for pos in range(every, doc.size - 2 * chunk, every):
        b0 = pos + (-pos) % chunk
        for b in range(b0, min(b0 + every - chunk, doc.size - chunk), chunk):
            region = doc[b - run - 1 : b + 1]
            if not ((region == 0x22) | (region == 0x5C) | (region >= 0x80)).any():
                doc[b - run : b] = 0x5C
                planted += 1
                break
    return planted
