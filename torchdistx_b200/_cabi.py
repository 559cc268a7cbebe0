"""ctypes view of the C ABI declared in ``include/tdx_init.h`` (libtdx_init.so).

The product path (``torchdistx_b200._C``) links the library directly; this module exists
so that tests, the kernel sweep and ``bench.py`` can drive the kernels *through the C ABI*
with raw device pointers, exactly as a non-Python host (the reference's C++ runtime,
``src/cc/torchdistx/deferred_init.cc``) would.

There is deliberately no fallback: if the shared library is missing, importing the
loader raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterable, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
# TDX_INIT_LIB: kernel-experiment builds only (benchmarks/); the product always uses the in-tree library
LIB_PATH = os.environ.get("TDX_INIT_LIB") or os.path.join(_HERE, "libtdx_init.so")

# enums of tdx_init.h
TDX_F32, TDX_BF16, TDX_F16, TDX_I64 = 0, 1, 2, 3
TDX_RAW8, TDX_RAW16, TDX_RAW32, TDX_RAW64 = 8, 9, 10, 11
TDX_SRC_CONST, TDX_SRC_UNIFORM, TDX_SRC_NORMAL, TDX_SRC_IOTA = 0, 1, 2, 3
TDX_ALGO_DEFAULT, TDX_ALGO_ICDF16, TDX_ALGO_BM32, TDX_ALGO_BM16 = 0, 1, 2, 3
TDX_ALGO_WIDE32 = 2
TDX_ALGO_R7 = 0x10
TDX_ALGO_NOLUT = 0x20
TDX_EPI_MUL, TDX_EPI_ADD, TDX_EPI_ERFINV, TDX_EPI_CLAMP, TDX_EPI_RPOW, TDX_EPI_RECIP = 1, 2, 3, 4, 5, 6
TDX_MAX_EPI = 4
TDX_EPI_NOROUND = 0x100
TDX_FLAG_SRC_NOROUND = 0x1

EXPORTED_SYMBOLS = (
    "tdx_init_workspace_bytes",
    "tdx_init_launch",
    "tdx_init_prepare",
    "tdx_init_submit",
    "tdx_plan_upload",
    "tdx_plan_launch",
    "tdx_last_launch_count",
    "tdx_last_upload_bytes",
    "tdx_elems_per_block",
    "tdx_abi_version",
    "tdx_last_error",
)


class TdxEpiStep(ctypes.Structure):
    _fields_ = [("op", ctypes.c_uint32), ("a", ctypes.c_float), ("b", ctypes.c_float)]


class TdxInitDesc(ctypes.Structure):
    _fields_ = [
        ("dst", ctypes.c_void_p),
        ("elem_begin", ctypes.c_uint64),
        ("elem_count", ctypes.c_uint64),
        ("philox_seed", ctypes.c_uint64),
        ("philox_offset", ctypes.c_uint64),
        ("p0", ctypes.c_double),
        ("p1", ctypes.c_double),
        ("fill_bits", ctypes.c_uint64 * 2),
        ("dtype", ctypes.c_uint8),
        ("src", ctypes.c_uint8),
        ("algo", ctypes.c_uint8),
        ("n_epi", ctypes.c_uint8),
        ("reserved", ctypes.c_uint32),
        ("epi", TdxEpiStep * TDX_MAX_EPI),
    ]


class TdxPlan(ctypes.Structure):
    _fields_ = [("opaque", ctypes.c_uint64 * 512)]


assert ctypes.sizeof(TdxInitDesc) == 128, ctypes.sizeof(TdxInitDesc)

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Loads libtdx_init.so (built by ``__graft_entry__.build()`` / ``build_kernels.sh``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA kernel library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "torchdistx_b200 has no CPU fallback for CUDA tensors."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.tdx_init_workspace_bytes.restype = ctypes.c_size_t
    lib.tdx_init_workspace_bytes.argtypes = [ctypes.c_int]
    lib.tdx_init_prepare.restype = ctypes.c_int
    lib.tdx_init_prepare.argtypes = [ctypes.POINTER(TdxInitDesc), ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]
    lib.tdx_init_submit.restype = ctypes.c_int
    lib.tdx_init_submit.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.tdx_init_launch.restype = ctypes.c_int
    lib.tdx_init_launch.argtypes = [
        ctypes.POINTER(TdxInitDesc), ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.tdx_plan_upload.restype = ctypes.c_int
    lib.tdx_plan_upload.argtypes = [
        ctypes.POINTER(TdxInitDesc), ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
        ctypes.POINTER(TdxPlan)]
    lib.tdx_plan_launch.restype = ctypes.c_int
    lib.tdx_plan_launch.argtypes = [ctypes.POINTER(TdxPlan), ctypes.c_void_p, ctypes.c_void_p]
    lib.tdx_last_launch_count.restype = ctypes.c_int
    lib.tdx_last_upload_bytes.restype = ctypes.c_size_t
    lib.tdx_elems_per_block.restype = ctypes.c_int
    lib.tdx_elems_per_block.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.tdx_abi_version.restype = ctypes.c_int
    lib.tdx_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libtdx_init: error {rc}: {load().tdx_last_error().decode()}")


def fill_pattern(value_bits: int, itemsize: int) -> Sequence[int]:
    """16-byte store pattern: the element's bits replicated."""
    b = int(value_bits).to_bytes(itemsize, "little") * (16 // itemsize)
    return int.from_bytes(b[:8], "little"), int.from_bytes(b[8:], "little")


def make_desc(dst: int, *, dtype: int, src: int, elem_count: int, elem_begin: int = 0,
              seed: int = 0, offset: int = 0, p0: float = 0.0, p1: float = 1.0, algo: int = 0,
              fill_bits: int = 0, fill_itemsize: int = 0,
              epi: Iterable[tuple] = (), flags: int = 0) -> TdxInitDesc:
    d = TdxInitDesc()
    d.dst = dst
    d.elem_begin = elem_begin
    d.elem_count = elem_count
    d.philox_seed = seed & (2**64 - 1)
    d.philox_offset = offset
    d.p0, d.p1 = p0, p1
    d.dtype, d.src, d.algo = dtype, src, algo
    d.reserved = flags
    if src == TDX_SRC_CONST:
        lo, hi = fill_pattern(fill_bits, fill_itemsize)
        d.fill_bits[0], d.fill_bits[1] = lo, hi
    steps = list(epi)
    d.n_epi = len(steps)
    for i, st in enumerate(steps):
        d.epi[i].op = st[0]
        d.epi[i].a = st[1] if len(st) > 1 else 0.0
        d.epi[i].b = st[2] if len(st) > 2 else 0.0
    return d


def launch(descs: Sequence[TdxInitDesc], workspace_ptr: int, workspace_bytes: int,
           stream: int = 0) -> int:
    """tdx_init_launch over a Python list of descriptors; returns the number of kernel launches."""
    lib = load()
    arr = (TdxInitDesc * len(descs))(*descs)
    check(lib.tdx_init_launch(arr, len(descs), workspace_ptr, workspace_bytes, stream))
    return lib.tdx_last_launch_count()


def prepare(descs: Sequence[TdxInitDesc]) -> int:
    """tdx_init_prepare: lays the plan out on the host; returns the exact workspace size it needs
    (0: nothing to launch).  Follow with `submit` on the same thread."""
    lib = load()
    arr = (TdxInitDesc * len(descs))(*descs)
    need = ctypes.c_size_t(0)
    check(lib.tdx_init_prepare(arr, len(descs), ctypes.byref(need)))
    return int(need.value)


def submit(workspace_ptr: int, workspace_bytes: int, stream: int = 0) -> int:
    """tdx_init_submit of the plan `prepare` built; returns the number of kernel launches."""
    lib = load()
    check(lib.tdx_init_submit(workspace_ptr, workspace_bytes, stream))
    return lib.tdx_last_launch_count()
