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


# ------------------------------------------------------------------------------------------------
# stream pipeline (stage1_stream.cuh): chunk summaries -> two-level span scan -> carry words -> flatten
# ------------------------------------------------------------------------------------------------
def _must_give_up(data: bytes, mis: int) -> bool:
    """The speculation fails exactly when the 32 bytes before some 2 KiB chunk boundary (aligned coordinates) cannot decide
    the carries: all backslashes (is the chunk's first byte escaped?), or a quote preceded by 31 backslashes (is that
    quote escaped, i.e. is the previous byte a scalar?)."""
    a = bytes(mis) + bytes(data)  # chunk 0 has no look-behind, and 2048 - 32 > mis: the window is always document bytes
    for b in range(2048, len(a), 2048):
        w = a[b - 32 : b]
        if w == b"\\" * 32 or w == b"\\" * 31 + b'"':
            return True
    return False


def _check_stream(data: bytes, mis: int, flags: int = 0):
    want = oracle.stage1(data, flags=flags, impl="fast" if len(data) > 20000 else "ref")
    err, n, nw, idx, u8, gave_up = emu.stage1_stream(data, mis=mis, flags=flags)
    if not data:
        assert err == want.error
        return False
    # mis < 16 and chunk boundaries are multiples of 2048 >= 2048: the look-behind always lies inside the document or
    # covers document bytes only
    assert gave_up == _must_give_up(data, mis)
    if gave_up:
        return True
    assert err == want.error
    assert n == want.n
    assert nw == want.n_written
    assert np.array_equal(idx, want.indexes)
    assert u8 == want.utf8_error
    return False


def test_stream_pipeline_adversarial_corpus():
    gave_up = 0
    for name, data in cases.adversarial_cases(tile_bytes=(2048, 4096, 16384)):
        for mis in (0, 5):
            try:
                gave_up += _check_stream(data, mis)
                _check_stream(data, mis, flags=1)
            except AssertionError as e:  # pragma: no cover
                raise AssertionError(f"case {name} mis={mis}") from e
    assert gave_up > 0, "the corpus must contain inputs that defeat the speculation (long backslash runs at chunk ends)"


def test_stream_pipeline_many_blocks():
    """More than one block of 4096 chunk summaries (> 8 MiB), with strings that stay open across block boundaries."""
    unit = b'{"a":[1,2,{"b":"' + b"x" * 3000 + b'\\"y"}],"c":"\xe2\x82\xac"},'
    data = b"[" + unit * ((20 << 20) // len(unit)) + b'"' + b"s" * (9 << 20) + b'",0]'
    assert len(data) > 3 * 4096 * 2048
    _check_stream(data, 0)
    _check_stream(data, 7)
    _check_stream(data[:-5], 3)  # unclosed


@settings(max_examples=150, deadline=None)
@given(st.lists(st.sampled_from(list(cases.NASTY)), min_size=1, max_size=9000), st.integers(0, 15))
def test_stream_pipeline_fuzz_nasty(xs, mis):
    _check_stream(bytes(xs), mis)


# ------------------------------------------------------------------------------------------------
# balanced flatten kernel (stage1_split.cuh: stage1_flatten2_kernel): its arithmetic and its steps for one unit
# ------------------------------------------------------------------------------------------------
def test_flatten_division_by_multiplication_is_exact():
    L = emu.lib()
    # the kernel divides x = (prefix count + q - 1) <= K + q - 1 by q, with K <= 32 q indexes in the unit (32-bit products)
    for q in range(1, 64):
        for x in range(0, min(4200, 34 * q + 2)):
            assert L.emu_fl2_div(x, q) == x // q, (x, q)


def test_flatten_drop_high_bits():
    rng = random.Random(11)
    L = emu.lib()
    for _ in range(4000):
        w = rng.getrandbits(32) & rng.getrandbits(32) if rng.random() < 0.5 else rng.getrandbits(32)
        if w == 0:
            continue
        bits = [i for i in range(31, -1, -1) if (w >> i) & 1]      # highest first
        for r in {0, len(bits) - 1, rng.randrange(len(bits))}:
            want = 0
            for i in bits[r:]:
                want |= 1 << i
            assert L.emu_drop_high_bits(w, r) == want, (hex(w), r)


def _unit_want(words, base):
    return np.array([base + 32 * k + i for k in range(128) for i in range(32) if (int(words[k]) >> i) & 1], dtype=np.uint32)


def test_flatten_unit_every_density():
    rng = np.random.default_rng(5)
    for trial in range(400):
        density = [0.0, 0.002, 0.01, 0.03, 0.08, 0.16, 0.1875, 0.3][trial % 8]
        bits = rng.random((128, 32)) < density
        if trial % 5 == 0:        # long stretches with no index at all (strings), also at both ends
            lo, hi = sorted(int(v) for v in rng.integers(0, 129, 2))
            bits[lo:hi] = False
        if trial % 7 == 0:        # a few full words next to empty ones
            for k in rng.integers(0, 128, 3):
                bits[int(k)] = True
        words = (bits * (1 << np.arange(32, dtype=np.uint64))).sum(axis=1).astype(np.uint32)
        want = _unit_want(words, 1000 * trial)
        rc, got = emu.flatten_unit(words, value_base=1000 * trial)
        if want.size > 768:
            assert rc == 1
        else:
            assert rc == 0, (trial, rc)
            assert np.array_equal(got, want), trial
    # the smallest and the largest units the balanced path takes
    for k, i in ((0, 0), (127, 31), (64, 5)):
        words = np.zeros(128, dtype=np.uint32)
        words[k] = 1 << i
        rc, got = emu.flatten_unit(words)
        assert rc == 0 and got.tolist() == [32 * k + i]
    words = np.zeros(128, dtype=np.uint32)
    words[:24] = 0xFFFFFFFF                                   # exactly 768 indexes in the first 24 words
    rc, got = emu.flatten_unit(words)
    assert rc == 0 and np.array_equal(got, np.arange(768, dtype=np.uint32))
