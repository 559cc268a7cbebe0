"""Builds torchdistx_b200/_C (the C++ recorder/planner extension) in-tree.

Links against the C-ABI kernel library torchdistx_b200/libtdx_init.so (build_kernels.sh).
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "torchdistx_b200")
SRC = os.path.join(PKG, "csrc", "recorder")


def build(verbose: bool = False) -> str:
    from torch.utils import cpp_extension

    build_dir = os.path.join(ROOT, "build", "_C")
    os.makedirs(build_dir, exist_ok=True)
    cpp_extension.load(
        name="_C",
        sources=[os.path.join(SRC, f) for f in ("fake_tensor.cc", "tape.cc", "planner.cc", "bindings.cc", "public_api.cc")],
        extra_include_paths=[os.path.join(ROOT, "include"), SRC, "/usr/local/cuda/include"],
        extra_cflags=["-O2", "-std=c++17", "-fvisibility=hidden"],
        extra_ldflags=[f"-L{PKG}", "-ltdx_init", '-Wl,-rpath,\'$$ORIGIN\'', f"-Wl,-rpath,{PKG}", "-lc10_cuda", "-Wl,--no-as-needed", "-ltorch_cuda", "-Wl,--as-needed"],
        build_directory=build_dir,
        with_cuda=False,
        verbose=verbose,
        is_python_module=False,
    )
    out = os.path.join(PKG, "_C.so")
    shutil.copyfile(os.path.join(build_dir, "_C.so"), out)
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
