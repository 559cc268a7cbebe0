"""ctypes loader of oracle/tdx_oracle.c (TEST INFRASTRUCTURE: tests/, smoke(), bench cpu_baseline)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libtdx_oracle.so")


def build() -> str:
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "tdx_oracle.c")):
            build()
        _lib = ctypes.CDLL(SO)
        _lib.tdx_oracle_generate.restype = ctypes.c_int
        _lib.tdx_oracle_generate.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _lib.tdx_oracle_philox4x32.restype = None
        _lib.tdx_oracle_philox4x32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib.tdx_oracle_offset_increment.restype = ctypes.c_uint64
        _lib.tdx_oracle_offset_increment.argtypes = [ctypes.c_uint64]
    return _lib


def philox4x32(ctr, key, rounds=10):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().tdx_oracle_philox4x32(c, k, rounds, o)
    return [int(x) for x in o]


def generate(desc) -> np.ndarray:
    """Expected content (raw element bits as a numpy array) of the buffer a descriptor writes.
    `desc` is a torchdistx_b200._cabi.TdxInitDesc (only its bytes are read)."""
    from torchdistx_b200 import _cabi as C

    isz = {C.TDX_F32: 4, C.TDX_RAW32: 4, C.TDX_RAW64: 8, C.TDX_I64: 8, C.TDX_RAW8: 1}.get(desc.dtype, 2)
    np_dtype = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[isz]
    out = np.empty(int(desc.elem_count), dtype=np_dtype)
    rc = lib().tdx_oracle_generate(ctypes.byref(desc), out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("descriptor is outside the oracle's specification")
    return out


def offset_increment(numel: int) -> int:
    return int(lib().tdx_oracle_offset_increment(numel))
