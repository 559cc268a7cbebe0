"""The C++ surface (include/torchdistx_b200/{fake,deferred_init}.h, exported by _C.so): a C++ caller
of the reference's installed headers (reference src/cc/torchdistx/fake.h:34-83,
deferred_init.h:25-37) finds the same names and behaviour.  A small C++ program is compiled against
the headers, linked with _C.so and run inside this process."""
import ctypes
import os
import subprocess
import sysconfig

import torch
from torch.utils import cpp_extension

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "torchdistx_b200")
SRC = os.path.join(ROOT, "tests", "cpp", "public_api_check.cc")
OUT_DIR = os.path.join(ROOT, "build", "tests")
OUT = os.path.join(OUT_DIR, "libpublic_api_check.so")


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [SRC, os.path.join(PKG, "_C.so"), os.path.join(ROOT, "include", "torchdistx_b200", "fake.h"),
            os.path.join(ROOT, "include", "torchdistx_b200", "deferred_init.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", SRC, "-o", OUT,
           f"-I{os.path.join(ROOT, 'include')}"]
    cmd += [f"-isystem{p}" for p in cpp_extension.include_paths()] + [f"-isystem{sysconfig.get_paths()['include']}"]
    cmd += [f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
            f"-L{PKG}", "-l:_C.so", f"-Wl,-rpath,{PKG}", f"-L{lib}", "-lc10", "-ltorch_cpu", "-ltorch", f"-Wl,-rpath,{lib}"]
    subprocess.run(cmd, check=True, timeout=600)
    return OUT


def test_exported_names_match_the_reference_headers():
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", os.path.join(PKG, "_C.so")], check=True,
                         capture_output=True, text=True).stdout
    for name in ("torchdistx::enterFakeMode(bool)", "torchdistx::leaveFakeMode()", "torchdistx::isFakeModeActive()",
                 "torchdistx::isFake(at::TensorBase const&)", "torchdistx::FakeTensor::FakeTensor(at::TensorBase const&, bool)",
                 "torchdistx::FakeTensor::toMeta() const", "torchdistx::FakeTensor::meta_storage() const",
                 "torchdistx::asFake(at::TensorBase const&)", "torchdistx::unsafeAsFake(at::TensorBase const&)",
                 "torchdistx::enterDeferredInit()", "torchdistx::leaveDeferredInit()",
                 "torchdistx::canMaterialize(at::Tensor const&)", "torchdistx::materializeTensor(at::Tensor const&)"):
        assert name in out, name


def test_cpp_caller_of_the_reference_headers():
    from torchdistx_b200 import _C  # noqa: F401  (the library the program links against, loaded first)

    lib = ctypes.CDLL(build())
    lib.tdx_public_api_check.restype = ctypes.c_int
    lib.tdx_public_api_check_error.restype = ctypes.c_char_p
    rc = lib.tdx_public_api_check()
    assert rc == 0, lib.tdx_public_api_check_error().decode()
