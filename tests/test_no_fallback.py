"""The product never routes through the oracle or a CPU fallback, and fails loudly without its
CUDA library."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def product_sources():
    for top in ("torchdistx_b200", "torchdistx", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cc", ".cu", ".cuh", ".h")):
                    yield os.path.join(dirpath, f)


def test_product_never_imports_or_links_the_oracle():
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#include\s+[\"<].*oracle)", re.M)
    offenders = [p for p in product_sources() if pat.search(open(p).read())]
    assert offenders == []
    # only tests/, __graft_entry__.py and bench.py may touch oracle/
    users = []
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "build", "gpurun_out", "oracle", "__pycache__", "profiles")]
        for f in files:
            if f.endswith(".py"):
                p = os.path.join(dirpath, f)
                if re.search(r"^\s*(from|import)\s+oracle\b", open(p).read(), re.M):
                    users.append(os.path.relpath(p, ROOT))
    allowed = ("tests/", "__graft_entry__.py", "bench.py")
    assert all(u.startswith(allowed) or u in allowed for u in users), users


def test_missing_kernel_library_is_an_import_error():
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['TDX_INIT_LIB'] = '/nonexistent/libtdx_init.so';"
            "from torchdistx_b200 import _cabi\n"
            "try:\n    _cabi.load()\nexcept ImportError as e:\n    print('IMPORT_ERROR', 'no CPU fallback' in str(e))") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    assert "IMPORT_ERROR True" in out


def test_native_module_depends_on_the_kernel_library():
    so = os.path.join(ROOT, "torchdistx_b200", "_C.so")
    dyn = subprocess.run(["readelf", "-d", so], capture_output=True, text=True, check=True).stdout
    assert "libtdx_init.so" in dyn  # DT_NEEDED: importing torchdistx_b200 without it fails


def test_cuda_tensors_are_never_materialised_on_the_cpu():
    """A program the planner cannot fuse is replayed by ATen on the recorded device; without a GPU
    that fails loudly instead of silently producing a CPU tensor."""
    import pytest
    import torch

    from torchdistx_b200.deferred_init import deferred_init, materialize_tensor
    from torchdistx_b200.fake import fake_mode

    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    from torchdistx_b200 import _C

    _C.enter_fake_mode(True)  # lets `device="cuda"` through without a GPU, as fake_mode(fake_cuda=True)
    try:
        p = deferred_init(lambda: torch.nn.Parameter(torch.empty(8, device="cuda").normal_()))
    finally:
        _C.leave_fake_mode()
    assert p.device.type == "cuda"
    with pytest.raises((RuntimeError, AssertionError)):
        materialize_tensor(p)
