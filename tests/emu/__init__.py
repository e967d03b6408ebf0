"""Builds and loads the host-side lane emulator (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "emulate_stage1.cpp")
_CORE = os.path.join(_HERE, "..", "..", "mojo_simdjson_b200", "csrc", "stage1_core.cuh")
_SO = os.path.join(_HERE, "libemu_stage1.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        newest = max(os.path.getmtime(_SRC), os.path.getmtime(_CORE))
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", _SO, _SRC])
        L = C.CDLL(_SO)
        L.emu_stage1.restype = C.c_int32
        L.emu_stage1.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64,
                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.c_uint32]
        L.emu_stage1_stream.restype = C.c_int32
        L.emu_stage1_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_uint32]
        L.emu_bitplanes32.restype = None
        L.emu_bitplanes32.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_fl2_div.restype = C.c_uint32
        L.emu_fl2_div.argtypes = [C.c_uint32, C.c_uint32]
        L.emu_drop_high_bits.restype = C.c_uint32
        L.emu_drop_high_bits.argtypes = [C.c_uint32, C.c_uint32]
        L.emu_flatten_unit.restype = C.c_int32
        L.emu_flatten_unit.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def stage1(data: bytes, mis: int = 0, warps: int = 8, flags: int = 0, cap=None):
    L = lib()
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    n_bytes = int(a.size)
    if cap is None:
        cap = n_bytes + 3
    out = np.full(max(cap, 1), 0xDEADBEEF, dtype=np.uint32)
    n = C.c_uint32(0xFFFFFFFF)
    nw = C.c_uint64(0)
    u8 = C.c_int32(0)
    err = L.emu_stage1(a.ctypes.data if n_bytes else None, n_bytes, mis, warps, out.ctypes.data, cap,
                       C.byref(n), C.byref(nw), C.byref(u8), flags)
    assigned = n.value != 0xFFFFFFFF
    keep = min(int(nw.value) + (3 if assigned else 0), cap)
    return err, (n.value if assigned else None), int(nw.value), out[:keep].copy(), int(u8.value)


def stage1_stream(data: bytes, mis: int = 0, flags: int = 0, cap=None):
    """The stream pipeline's arithmetic on the host.  Returns (err, n, n_written, out, utf8, gave_up)."""
    L = lib()
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    n_bytes = int(a.size)
    if cap is None:
        cap = n_bytes + 3
    out = np.full(max(cap, 1), 0xDEADBEEF, dtype=np.uint32)
    n = C.c_uint32(0xFFFFFFFF)
    nw = C.c_uint64(0)
    u8 = C.c_int32(0)
    spec = C.c_int32(0)
    err = L.emu_stage1_stream(a.ctypes.data if n_bytes else None, n_bytes, mis, out.ctypes.data, cap,
                              C.byref(n), C.byref(nw), C.byref(u8), C.byref(spec), flags)
    assigned = n.value != 0xFFFFFFFF
    keep = min(int(nw.value) + (3 if assigned else 0), cap)
    return err, (n.value if assigned else None), int(nw.value), out[:keep].copy(), int(u8.value), bool(spec.value)


def flatten_unit(words: np.ndarray, value_base: int = 0, cap: int = 768):
    """The balanced flatten kernel's steps for one unit of 128 mask words.  Returns (rc, indexes)."""
    L = lib()
    w = np.ascontiguousarray(words, dtype=np.uint32)
    assert w.size == 128
    out = np.zeros(4096, dtype=np.uint32)
    count = C.c_uint32(0)
    rc = L.emu_flatten_unit(w.ctypes.data, value_base, cap, out.ctypes.data, C.byref(count))
    return rc, out[: count.value].copy()
