"""Python surface over the compiled reference runtime (oracle/_ref/_tdx_ref.so).

TEST INFRASTRUCTURE ONLY -- never imported by torchdistx_b200.

Restates, function for function, the reference's pure-Python layer so that tests can call the
reference exactly like a user would:
  fake_mode / is_fake / meta_like     <- reference src/python/torchdistx/fake.py:43-82
  deferred_init / is_deferred /
  materialize_tensor / materialize_module <- reference src/python/torchdistx/deferred_init.py:19-124

IMPORTANT: the reference runtime registers boxed fallbacks on DispatchKey::Fake and
DispatchKey::DeferredInit (reference fake.cc:546-548, deferred_init.cc:880-883); so does
torchdistx_b200._C.  The two therefore cannot live in one process: use oracle/ref_driver.py
(a subprocess) from tests that also import torchdistx_b200.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
from contextlib import contextmanager

# torch.distributed.fsdp probes `import torchdistx` ($TORCH/distributed/fsdp/_init_utils.py:53-57);
# in this repo that name resolves to the shim over torchdistx_b200, whose native module would
# register the same dispatch-key fallbacks as the reference.  Block it in oracle processes.
if "torchdistx_b200._C" in sys.modules:
    raise ImportError("the reference oracle cannot share a process with torchdistx_b200")
sys.modules.setdefault("torchdistx", None)  # type: ignore[arg-type]
sys.modules.setdefault("torchdistx_b200", None)  # type: ignore[arg-type]

import torch  # noqa: E402

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "_tdx_ref.so")


def _load():
    if not os.path.exists(_SO):
        from . import build_ref  # type: ignore
        build_ref.build()
    loader = importlib.machinery.ExtensionFileLoader("_tdx_ref", _SO)
    spec = importlib.util.spec_from_loader("_tdx_ref", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


_C = _load()


@contextmanager
def fake_mode(*, fake_cuda: bool = False):
    _C.enter_fake_mode(fake_cuda)
    try:
        yield
    finally:
        _C.leave_fake_mode()


def is_fake(t):
    return _C.is_fake(t)


def meta_like(t):
    return _C.meta_like(t)


def deferred_init(module_fn, *args, **kwargs):
    _C.enter_deferred_init()
    try:
        return module_fn(*args, **kwargs)
    finally:
        _C.leave_deferred_init()


def is_deferred(obj):
    if isinstance(obj, torch.Tensor):
        return _C.can_materialize(obj)
    if isinstance(obj, torch.nn.Module):
        return any(_C.can_materialize(t) for t in list(obj.parameters()) + list(obj.buffers()))
    raise ValueError("`obj` must be of type `Tensor` or `Module`.")


def materialize_tensor(t):
    """The oracle binding returns a plain at::Tensor; re-wrap Parameters like the reference does
    (reference _C/deferred_init.cc:94 makeVariable(Py_TYPE(var), ...))."""
    out = _C.materialize_tensor(t)
    if out is not t and isinstance(t, torch.nn.Parameter) and not isinstance(out, torch.nn.Parameter):
        out = torch.nn.Parameter(out, requires_grad=t.requires_grad)
    return out


def materialize_module(module, buffers_only: bool = False, check_fn=None) -> None:
    # children first, then own parameters, then own buffers (reference deferred_init.py:104-124)
    for child in module.children():
        materialize_module(child, buffers_only, check_fn)
    if check_fn is not None and not check_fn(module):
        return
    groups = ([] if buffers_only else [module._parameters]) + [module._buffers]
    for group in groups:
        for key, t in group.items():
            if t is not None:
                group[key] = materialize_tensor(t)
