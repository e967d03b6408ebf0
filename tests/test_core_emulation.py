"""CPU checks of the kernel's arithmetic (stage1_core.cuh) through the host-side lane emulator.

The emulator compiles the same header the CUDA kernel compiles, so a pass here means the bit tricks are
right; what remains GPU-only is memory movement, warp collectives and the look-back protocol.
"""
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle
from tests import cases, emu


def test_bitplanes_match_definition():
    rng = random.Random(3)
    L = emu.lib()
    for _ in range(200):
        b = np.frombuffer(bytes(rng.randrange(256) for _ in range(32)), dtype=np.uint8).copy()
        planes = np.zeros(8, dtype=np.uint32)
        L.emu_bitplanes32(b.ctypes.data, planes.ctypes.data)
        for k in range(8):
            want = 0
            for i in range(32):
                want |= ((int(b[i]) >> k) & 1) << i
            assert int(planes[k]) == want


def _check(data: bytes, mis: int, warps: int, flags: int = 0):
    want = oracle.stage1(data, flags=flags, impl="fast" if len(data) > 20000 else "ref")
    err, n, nw, idx, u8 = emu.stage1(data, mis=mis, warps=warps, flags=flags)
    assert err == want.error
    assert n == want.n
    assert nw == want.n_written
    assert np.array_equal(idx, want.indexes)
    assert u8 == want.utf8_error


@pytest.mark.parametrize("warps", [2, 8, 32])
def test_adversarial_corpus(warps):
    tiles = tuple(sorted({warps * 2048, 4096}))
    for name, data in cases.adversarial_cases(tile_bytes=tiles):
        for mis in (0, 5):
            try:
                _check(data, mis, warps)
                _check(data, mis, warps, flags=1)
            except AssertionError as e:  # pragma: no cover
                raise AssertionError(f"case {name} mis={mis} warps={warps}") from e


def test_every_misalignment():
    data = b'{"k":"v\\"x","a":[1,2,3],"u":"\xe2\x82\xac"}' * 40
    for mis in range(16):
        _check(data, mis, 2)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.sampled_from(list(cases.NASTY)), min_size=1, max_size=9000), st.integers(0, 15))
def test_fuzz_nasty(xs, mis):
    _check(bytes(xs), mis, 2, flags=1)


@settings(max_examples=200, deadline=None)
@given(st.binary(min_size=1, max_size=5000), st.integers(0, 15))
def test_fuzz_binary(data, mis):
    _check(data, mis, 2, flags=1)
