"""Pins the CPU oracle: reference fixtures, known answers, cross-formulation agreement.

These run without a GPU (-m "not gpu").
"""
import itertools
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle
from tests import cases


@pytest.mark.parametrize("fx", cases.golden_fixtures(), ids=lambda f: f["file"])
@pytest.mark.parametrize("impl", ["ref", "spec", "fast"])
def test_reference_fixtures(fx, impl):
    """tests/test_stage_1.mojo:85-96 + :43-82, restated: error 0, exact mask equality, trailer."""
    data = fx["input"].encode("utf-8")
    r = oracle.stage1(data, impl=impl)
    assert r.error == 0
    got = [int(x) for x in r.indexes[: r.n]]
    assert all(a < b for a, b in zip(got, got[1:]))
    mask = [" "] * len(fx["mask"])
    for i in got:
        mask[i] = "1"
    mask = "".join(mask)
    if fx["harness_negative"]:
        # the reference's harness self-tests: the file's line 2 is deliberately wrong
        assert mask != fx["mask"]
        assert got == [i for i in fx["claimed_indexes"] if i != 6]
    else:
        assert mask == fx["mask"]
        assert got == fx["claimed_indexes"]
    n = r.n
    assert list(r.indexes[n : n + 3]) == [len(data), len(data), 0]


def test_fixture_structurals_sit_on_structural_characters():
    # assert_tagging_is_correct, tests/test_stage_1.mojo:28-40
    for fx in cases.golden_fixtures():
        if fx["harness_negative"]:
            continue
        for i in fx["claimed_indexes"]:
            assert fx["input"][i] in '{}[]:,tfn-0123456789"'


@pytest.mark.parametrize("case", cases.KNOWN_ANSWERS, ids=lambda c: repr(c)[:24])
@pytest.mark.parametrize("impl", ["ref", "spec", "fast"])
def test_known_answers(case, impl):
    data, err, n, idx = case
    r = oracle.stage1(data, impl=impl)
    assert r.error == err
    assert r.n == n
    if idx is not None:
        assert [int(x) for x in r.indexes[: len(idx)]] == idx
        assert r.n_written == len(idx)
    if n is not None:
        assert list(r.indexes[n : n + 3]) == [len(data), len(data), 0]


def test_utf8_flag_priority():
    # the UTF-8 verdict is the last slot (json_structural_indexer.mojo:185-186): it never masks
    # UNCLOSED_STRING / UNESCAPED_CHARS / EMPTY, and it is off by default (reference: stub)
    bad = b'["\xc0\x80"]'
    assert oracle.stage1(bad).error == 0
    assert oracle.stage1(bad).utf8_error == 1
    assert oracle.stage1(bad, flags=oracle.FLAG_VALIDATE_UTF8).error == oracle.UTF8_ERROR
    assert oracle.stage1(b'"\xff', flags=oracle.FLAG_VALIDATE_UTF8).error == oracle.UNCLOSED_STRING
    assert oracle.stage1(b'"\xff\x01"', flags=oracle.FLAG_VALIDATE_UTF8).error == oracle.UNESCAPED_CHARS
    assert oracle.stage1(b" \xff ", flags=oracle.FLAG_VALIDATE_UTF8).error == oracle.UTF8_ERROR


def test_capacity_rule():
    data = b"[1,2,3]"
    assert oracle.stage1(data, cap=10).error == 0
    assert oracle.stage1(data, cap=9).error == oracle.CAPACITY
    assert oracle.stage1(data, impl="spec", cap=9).error == oracle.CAPACITY


def test_classifier_all_bytes_both_shuffle_semantics():
    """ws / op sets for all 256 byte values, under both plausible _dynamic_shuffle behaviours."""
    L = oracle.lib()
    want_ws = {0x20, 0x09, 0x0A, 0x0D}
    want_op = {0x2C, 0x3A, 0x5B, 0x5D, 0x7B, 0x7D, 0x0C, 0x1A}
    try:
        for variant in (0, 1):
            L.oracle_set_shuffle_variant(variant)
            for b in range(256):
                ref = oracle.stage1(bytes([b]) + b"1 ", impl="ref")
                spec = oracle.stage1(bytes([b]) + b"1 ", impl="spec")
                assert np.array_equal(ref.indexes, spec.indexes) and ref.error == spec.error
                got = [int(x) for x in ref.indexes[: ref.n_written]]
                if b in want_ws:
                    assert got == [1]
                elif b in want_op:
                    assert got == [0, 1]
                elif b == 0x22:
                    assert got == [0]  # opens a string that swallows the rest
                else:
                    assert got == [0]  # scalar glued to the following scalar
    finally:
        L.oracle_set_shuffle_variant(0)


def _agree(data: bytes):
    a = oracle.stage1(data, impl="ref")
    b = oracle.stage1(data, impl="spec")
    c = oracle.stage1(data, impl="fast")
    for o in (b, c):
        assert o.error == a.error
        assert o.n == a.n
        assert o.n_written == a.n_written
        assert np.array_equal(o.indexes, a.indexes)
    assert a.utf8_error == b.utf8_error  # DFA vs Keiser-Lemire tables
    try:
        data.decode("utf-8")
        ok = True
    except UnicodeDecodeError:
        ok = False
    assert a.utf8_error == (0 if ok else 1)


def test_formulations_agree_on_adversarial_corpus():
    for name, data in cases.adversarial_cases():
        try:
            _agree(data)
        except AssertionError as e:  # pragma: no cover
            raise AssertionError(f"case {name}") from e


@settings(max_examples=300, deadline=None)
@given(st.binary(min_size=0, max_size=600))
def test_formulations_agree_fuzz_binary(data):
    _agree(data)


@settings(max_examples=300, deadline=None)
@given(st.lists(st.sampled_from(list(cases.NASTY)), min_size=0, max_size=2000))
def test_formulations_agree_fuzz_nasty(xs):
    _agree(bytes(xs))


def _py_valid(b: bytes) -> bool:
    try:
        b.decode("utf-8")
        return True
    except UnicodeDecodeError:
        return False


def test_utf8_validators_agree_exhaustive_short():
    interesting = [0x00, 0x41, 0x7F, 0x80, 0x8F, 0x90, 0x9F, 0xA0, 0xBF, 0xC0, 0xC1, 0xC2, 0xDF, 0xE0, 0xE1, 0xEC, 0xED,
                   0xEE, 0xEF, 0xF0, 0xF1, 0xF3, 0xF4, 0xF5, 0xF7, 0xF8, 0xFF]
    for n in (1, 2, 3):
        for tup in itertools.product(interesting, repeat=n):
            b = bytes(tup)
            ok = _py_valid(b)
            assert oracle.utf8_valid(b, "dfa") == ok, b
            assert oracle.utf8_valid(b, "kl") == ok, b
    rng = random.Random(7)
    for _ in range(20000):
        b = bytes(rng.choice(interesting) for _ in range(rng.randint(4, 9)))
        ok = _py_valid(b)
        assert oracle.utf8_valid(b, "dfa") == ok, b
        assert oracle.utf8_valid(b, "kl") == ok, b


def test_digest_is_order_sensitive():
    a = np.arange(10, dtype=np.uint32)
    b = a.copy()
    b[[2, 3]] = b[[3, 2]]
    assert oracle.index_digest(a) != oracle.index_digest(b)
    assert oracle.index_digest(a) == oracle.index_digest(a.copy())


def test_structural_bytes_side_output():
    """Every structural index of a valid document points at an operator, a quote or the first byte of a scalar."""
    data = b'{"a":[1,-2.5e3,true,null,"x\\"y"],"b":{}}'
    r = oracle.stage1(data)
    got = oracle.structural_bytes(data, r.indexes[: r.n])
    assert bytes(got) == b'{":[1,-,t,n,"],":{}}'
    assert list(oracle.structural_bytes(data, r.indexes[: r.n + 3])[-3:]) == [0, 0, ord("{")]  # len, len, 0


def test_document_starts_side_output():
    data = b'{"a":[1,2]}\n[3]\n4 "s"\n{"b":{}}'
    r = oracle.stage1(data)
    sb = oracle.structural_bytes(data, r.indexes[: r.n])
    assert bytes(sb) == b'{":[1,2]}[3]4"{":{}}'
    starts = oracle.document_starts(sb)
    assert [k for k, f in enumerate(starts) if f] == [0, 9, 12, 13, 14]   # { [ 4 " {


# ------------------------------------------------------------------------------------------------
# "cpu_simd" (oracle/stage1_simd.c): the AVX-512 / AVX2 + pclmulqdq baseline bench.py times beside the faithful port
# ------------------------------------------------------------------------------------------------
def _simd_levels():
    try:
        have = oracle.simd_level()
    except Exception:
        return []
    return [lvl for lvl in (1, 2) if lvl <= have]


@pytest.mark.parametrize("level", [1, 2])
def test_cpu_simd_baseline_matches_the_oracle(level):
    if level not in _simd_levels():
        pytest.skip("this CPU lacks the instruction set")
    for name, data in cases.adversarial_cases():
        for flags in (0, 1):
            w = oracle.stage1(data, flags=flags, impl="fast" if len(data) > 20000 else "ref")
            g = oracle.stage1_simd(data, flags=flags, level=level)
            assert (g.error, g.n, g.n_written, g.utf8_error) == (w.error, w.n, w.n_written, w.utf8_error), (name, flags)
            assert np.array_equal(g.indexes[: w.indexes.size], w.indexes), (name, flags)


@pytest.mark.parametrize("level", [1, 2])
def test_cpu_simd_baseline_capacity(level):
    if level not in _simd_levels():
        pytest.skip("this CPU lacks the instruction set")
    data = b"[" + b"1," * 3000 + b"1]"
    w = oracle.stage1(data)
    for cap in (w.n + 3, w.n + 2, 100, 3):
        g = oracle.stage1_simd(data, cap=cap, level=level)
        if cap >= w.n + 3:
            assert g.error == 0 and np.array_equal(g.indexes, w.indexes)
        else:
            assert g.error == oracle.CAPACITY and g.n is None
            keep = min(cap, w.n)
            assert np.array_equal(g.indexes[:keep], w.indexes[:keep])


def test_cpu_simd_baseline_on_the_bench_workload():
    if not _simd_levels():
        pytest.skip("this CPU lacks the instruction set")
    from mojo_simdjson_b200 import synth

    for doc in (synth.status_array(3 << 20), synth.ndjson(2 << 20), synth.twitter_like()):
        w = oracle.stage1(doc, impl="fast")
        g = oracle.stage1_simd(doc)
        assert g.error == w.error == 0 and g.n == w.n and g.utf8_error == w.utf8_error == 0
        assert np.array_equal(g.indexes[: w.n + 3], w.indexes)
