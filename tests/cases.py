"""Shared input corpus for the oracle tests (CPU) and the GPU parity tests.

Every case is (name, bytes).  Sizes stay small enough for the oracle to finish in seconds.
The sets follow SURVEY.md section 8(c)/(d) config 5: carries x alignments, adversarial escapes,
quote density, UTF-8 edge cases, control bytes, the 0x0C/0x1A quirk, length sweeps.
"""
from __future__ import annotations

import json
import os
import random

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage1_fixtures.json")


def golden_fixtures():
    with open(GOLDEN, "r", encoding="utf-8") as f:
        return json.load(f)["fixtures"]


# (input, expected error, expected n or None, expected indexes or None) -- SURVEY.md section 8(c)
KNOWN_ANSWERS = [
    (b"", 13, None, None),
    (b" " * 10, 13, 0, []),
    (b" " * 128, 13, 0, []),
    (b'"abc', 15, None, [0]),
    (b'"a\x01b"', 14, None, [0]),
    (b'"a\nb"', 14, None, [0]),
    (b"[\x01]", 0, 3, [0, 1, 2]),
    (b"[\x0c]", 0, 3, [0, 1, 2]),
    (b"[\x1a]", 0, 3, [0, 1, 2]),
    (b'"\xff\xfe"', 0, 1, [0]),
    (b'["\xc0\x80"]', 0, 3, [0, 1, 5]),
    (b"[" + b"1," * 31 + b"]", 0, 64, list(range(64))),
    (b"[" + b"1," * 63 + b"]", 0, 128, list(range(128))),
    (b"[" + b"1," * 127 + b"]", 0, 256, list(range(256))),
    (b'"' + b"a" * 61 + b'\\\\\\"x"', 0, 1, [0]),
    (b'"' + b"a" * 62 + b'\\"x"', 0, 1, [0]),
    (b'"' + b"a" * 126 + b'\\"x"', 0, 1, [0]),
    (b"[" + b" " * 62 + b"12345]", 0, 3, [0, 63, 68]),
    (b'{"k":"' + b"x" * 200 + b'"}', 0, 5, [0, 1, 4, 5, 207]),
    (b'"' + b"\\" * 130 + b'"', 0, 1, [0]),
    (b'"' + b"\\" * 131 + b'"', 15, None, [0]),
    (b"[1] x y", 0, 5, [0, 1, 2, 4, 6]),
    (b'1"a"', 0, 1, [0]),
    (b'"a"1', 0, 2, [0, 3]),
    (b'{"a":1}\n{"b":2}\n[3]\n', 0, 13, [0, 1, 4, 5, 6, 8, 9, 12, 13, 14, 16, 17, 18]),
]

# valid extremes: U+0080, U+07FF, U+0800, U+FFFF, U+10000, U+10FFFF and some text
UTF8_VALID = [
    b"",
    "\u0080".encode("utf-8"),
    "\u07ff".encode("utf-8"),
    "\u0800".encode("utf-8"),
    "\uffff".encode("utf-8"),
    "\U00010000".encode("utf-8"),
    "\U0010ffff".encode("utf-8"),
    "h\u00e9llo".encode("utf-8"),
    "\u65e5\u672c\u8a9e".encode("utf-8"),
    "\U0001f600".encode("utf-8"),
]
UTF8_INVALID = [
    b"\xc0\x80", b"\xc1\xbf", b"\xe0\x80\x80", b"\xe0\x9f\xbf", b"\xf0\x80\x80\x80", b"\xf0\x8f\xbf\xbf",
    b"\xed\xa0\x80", b"\xed\xbf\xbf", b"\xf4\x90\x80\x80", b"\xf5\x80\x80\x80", b"\xf8\x88\x80\x80\x80", b"\xff", b"\xfe",
    b"\x80", b"\xbf", b"\xc2", b"\xe2\x82", b"\xf0\x9f\x98", b"\xc2\x41", b"\xe2\x41\x80", b"\xe2\x82\x41", b"\xf0\x41\x80\x80",
    b"\xf0\x9f\x41\x80", b"\xf0\x9f\x98\x41", b"\xc2\x80\x80", b"\xe2\x82\xac\x80", b"\xf0\x9f\x98\x80\x80",
]


def rng_bytes(rng: random.Random, n: int, alphabet: bytes) -> bytes:
    return bytes(rng.choice(alphabet) for _ in range(n))


JSONISH = b'"\\{}[]:, \n\tabcxyz0123456789-.eE' + b"tfn"
NASTY = JSONISH + bytes(range(0x00, 0x20)) + bytes([0x7F, 0x80, 0xBF, 0xC2, 0xE0, 0xED, 0xF0, 0xF4, 0xFF])


def adversarial_cases(tile_bytes=(4096, 8192, 16384), heavy: bool = False):
    """Deterministic adversarial inputs.  tile_bytes: tile sizes whose edges we want straddled."""
    cases = []

    def add(name, b):
        cases.append((name, bytes(b)))

    for fx in golden_fixtures():
        add(f"golden:{fx['file']}", fx["input"].encode("utf-8"))
    for i, (inp, *_rest) in enumerate(KNOWN_ANSWERS):
        add(f"known:{i}", inp)

    # (v) length sweep around every block edge that exists anywhere in either implementation
    rng = random.Random(1234)
    edges = sorted({1, 2, 3, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 2047, 2048, 2049}
                   | {t + d for t in tile_bytes for d in (-65, -64, -1, 0, 1, 63, 64, 65)})
    for n in edges:
        add(f"len:{n}:jsonish", rng_bytes(rng, n, JSONISH))
        add(f"len:{n}:digits", (b"[" + b"1," * n)[:n])
        add(f"len:{n}:string", (b'"' + b"a" * n)[: n - 1] + b'"' if n >= 2 else b"1")

    # truncated / complete multi-byte sequences at the very end of documents whose length is an exact multiple of every
    # chunk and tile size (the last lane of a FULL last chunk owns the end-of-input check), also one byte either side
    for total in sorted({2048, 4096, 6144} | set(tile_bytes) | {2 * t for t in tile_bytes}):
        for tail in (b"\xc3", b"\xe2\x82", b"\xf0\x9f\x98", "é".encode(), "€".encode(), "😀".encode()):
            for d in (-1, 0, 1):
                n = total + d
                add(f"utf8-end:{n}:{tail.hex()}", b'"' + b"a" * (n - 1 - len(tail)) + tail)
                add(f"utf8-end-closed:{n}:{tail.hex()}", b'"' + b"a" * (n - 2 - len(tail)) + tail + b'"')

    # (i) backslash runs of every length at several offsets, inside a string, then a quote
    long_runs = [4095, 4096, 4097] + ([65535, 65536, (1 << 20) + 1] if heavy else [])
    for run in list(range(1, 131)) + long_runs:
        for off in ([0, 1, 31, 62, 63, 64, 65, 127] if run > 8 else range(0, 130, 3)):
            body = b'["' + b"a" * off + b"\\" * run + b'"x", "y"]'
            add(f"bsrun:{run}@{off}", body)
    # runs ending exactly at / straddling tile edges
    for t in tile_bytes:
        for run in (1, 2, 3, 63, 64, 65, 128, 129):
            for end_delta in (-2, -1, 0, 1, 2):
                pre = t + end_delta - run - 2
                if pre < 0:
                    continue
                add(f"bsrun-tile:{t}:{run}:{end_delta}", b'["' + b"b" * pre + b"\\" * run + b'"q",1,"z"]')
    # whole tiles of backslashes (the all-backslash carry monoid)
    for t in tile_bytes[:2]:
        for extra in (0, 1, 2, 3):
            add(f"bs-alltile:{t}:{extra}", b'"' + b"\\" * (2 * t + extra) + b'" 1')
            add(f"bs-alltile-aligned:{t}:{extra}", b'"' + b"a" * (t - 1) + b"\\" * (t + extra) + b'" 1')

    # (ii) quote-dense / escape-dense
    add("quotes:dense", b'""' * 3000)
    add("quotes:odd", b'"' * 4097)
    add("esc:dense", b'"' + b'\\"' * 5000 + b'"')
    add("esc:dense2", b'["' + b'\\\\\\"' * 3000 + b'"]')
    add("mixed:strings", (b'{"k\\"ey":"va\\\\","x":[1,2,{"y":"\\\\\\""}]}' * 700))

    # strings and scalars spanning tile edges
    for t in tile_bytes:
        add(f"string-span:{t}", b'{"a":"' + b"s" * (2 * t + 17) + b'","b":123}')
        add(f"scalar-span:{t}", b"[" + b"7" * (t - 1) + b"," + b"8" * (t + 5) + b"]")
        add(f"ws-span:{t}", b"[" + b" " * (2 * t) + b"1]")
        for d in (-2, -1, 0, 1):
            add(f"quote-at-edge:{t}:{d}", b"[" + b" " * (t + d - 1) + b'"abc",' + b" " * 5 + b"12]")
            add(f"scalar-at-edge:{t}:{d}", b"[" + b"1" * (t + d - 1) + b' "q" ]')

    # (iv) control bytes in and out of strings, the 0x0C / 0x1A quirk
    for c in list(range(0x00, 0x20)) + [0x7F]:
        add(f"ctl-out:{c:02x}", b"[1," + bytes([c]) + b",2]")
        add(f"ctl-in:{c:02x}", b'["a' + bytes([c]) + b'b"]')
    add("ctl-in-far", b'["' + b"a" * 5000 + b"\x01" + b"a" * 5000 + b'"]')

    # (iii) UTF-8: every valid extreme and invalid form at offsets around 16B / 64B / tile edges
    seqs = UTF8_VALID + UTF8_INVALID
    for si, s in enumerate(seqs):
        for edge in (16, 64, 2048) + tuple(tile_bytes[:1]):
            for d in range(-4, 2):
                pre = edge + d - 2
                if pre < 0:
                    continue
                add(f"utf8:{si}@{edge}{d:+d}", b'["' + b"a" * pre + s + b'"]')
        add(f"utf8-eof:{si}", b'"' + s)          # sequence right at EOF
        add(f"utf8-bare:{si}", s)
        add(f"utf8-eof64:{si}", (b" " * 64)[: 64 - len(s)] + s)
        add(f"utf8-eof128:{si}", (b" " * 128)[: 128 - len(s)] + s)

    # random jsonish + nasty alphabets, several sizes
    rng = random.Random(0)
    for k in range(200 if heavy else 60):
        n = rng.choice([1, 7, 63, 64, 100, 500, 3000, 5000, 20000, 70000])
        add(f"rand-jsonish:{k}", rng_bytes(rng, n, JSONISH))
        add(f"rand-nasty:{k}", rng_bytes(rng, n, NASTY))
    return cases
