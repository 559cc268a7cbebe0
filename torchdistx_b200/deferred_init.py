"""Deferred module initialisation with a Blackwell-native materialiser.

Drop-in for ``torchdistx.deferred_init`` (reference src/python/torchdistx/deferred_init.py:19-124):
``deferred_init``, ``is_deferred``, ``materialize_tensor`` and ``materialize_module`` keep their
signatures, traversal order, return values and errors.

What differs is underneath ``materialize_*``.  The reference calls ``_C.materialize_tensor`` once
per tensor and each call replays the recorded aten ops one by one through the dispatcher
(reference deferred_init.py:104-113 -> deferred_init.cc:506-528).  Here ``materialize_module``
collects every tensor of the module tree (same order: children first, then the module's own
parameters, then its buffers) and hands the whole list to the native engine, which folds each
tensor's recorded program into one fused descriptor and runs all of them with a handful of
sm_100a kernel launches (``include/tdx_init.h``).  The extra keyword-only arguments select the
target device and dim-0 sharding:

    materialize_module(m, device="cuda")                       # recorded on cpu, built on the GPU
    materialize_module(m, device="cuda", shard=(rank, world))  # each rank builds only its rows
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple, TypeVar, Union

import torch
from torch import Tensor
from torch.nn import Module

from . import _C
from . import fake  # noqa: F401  (installs the fake-aware Tensor.__repr__)

__all__ = [
    "deferred_init",
    "is_deferred",
    "materialize_tensor",
    "materialize_module",
    "last_materialize_stats",
    "last_descriptors",
    "plan_report",
]

M = TypeVar("M", bound=Module)
DeviceLike = Union[str, torch.device, None]
Shard = Optional[Tuple[int, int]]


def deferred_init(module_fn: Callable[..., M], *args, **kwargs) -> M:
    """Runs ``module_fn(*args, **kwargs)`` with every tensor fake and every operation recorded.

    The returned module owns no memory; build it later with :func:`materialize_module` or
    :func:`materialize_tensor`.  Operations performed on the module after this function returns
    are not recorded.
    """
    _C.enter_deferred_init()
    try:
        return module_fn(*args, **kwargs)
    finally:
        _C.leave_deferred_init()


def is_deferred(obj: Union[Tensor, Module]) -> bool:
    """``True`` if the tensor -- or any parameter or buffer of the module -- still awaits
    materialisation."""
    if isinstance(obj, Tensor):
        return _C.can_materialize(obj)
    if isinstance(obj, Module):
        return any(_C.can_materialize(t) for t in obj.parameters()) or any(
            _C.can_materialize(t) for t in obj.buffers()
        )
    raise ValueError("`obj` must be of type `Tensor` or `Module`.")


def _device(device: DeviceLike) -> Optional[torch.device]:
    return None if device is None else torch.device(device)


def materialize_tensor(tensor: Tensor, *, device: DeviceLike = None, shard: Shard = None) -> Tensor:
    """Materialises ``tensor``; a real tensor is returned unchanged.

    Repeated calls return the same tensor object (the recording keeps a reference to it, so drop
    the fake tensor when it is no longer needed).
    """
    return _C.materialize_tensor(tensor, _device(device), shard)


def materialize_module(
    module: Module,
    buffers_only: bool = False,
    check_fn: Optional[Callable[[Module], bool]] = None,
    *,
    device: DeviceLike = None,
    shard: Shard = None,
) -> None:
    """Materialises ``module`` and its descendants in place.

    Args:
        module: the module to materialise.
        buffers_only: materialise buffers only, leave parameters fake.
        check_fn: called with every (sub)module; modules for which it returns ``False`` are
            skipped (their children are still visited).
        device: build the tensors on this device instead of the recorded one.
        shard: ``(rank, world_size)``: build only this rank's ``torch.chunk(..., dim=0)`` slice of
            every parameter (buffers and 0-dim tensors are replicated).  All ranks must hold the
            same generator state; see :func:`torchdistx_b200.parallel.sync_rng`.
    """
    # (a ValueError raised while a tensor is materialised comes back as the reference words it,
    # deferred_init.py:110-113: "'<key>' has already been materialized."; others pass unchanged)
    _C.materialize_module(module, buffers_only, check_fn, _device(device), shard)


def plan_report(module: Module) -> Dict[str, Dict[str, object]]:
    """What the planner will do with every parameter / buffer of a deferred module, without
    allocating or launching anything (works on a machine without a GPU): per tensor name,
    ``{"fusible", "source" (uninit|const|uniform|normal|opaque|real|materialized), "dtype", "numel",
    "p0", "p1", "n_epilogue", "rng_ops", "first_unfusable_op"}``."""
    out: Dict[str, Dict[str, object]] = {}
    for name, t in list(module.named_parameters()) + list(module.named_buffers()):
        out[name] = dict(_C.plan_info(t))
    return out


def last_descriptors():
    """The `TdxInitDesc` table the last ``materialize_*`` call on this thread submitted to the
    kernels, as a ctypes array (see ``include/tdx_init.h``)."""
    from . import _cabi

    raw = _C.last_descriptors()
    n = len(raw) // 128
    return (_cabi.TdxInitDesc * n).from_buffer_copy(raw)


def last_materialize_stats() -> Dict[str, int]:
    """Counters of the last ``materialize_*`` call on this thread (fused tensors, generic ops
    replayed, dead RNG passes elided, kernel launches, bytes written)."""
    return dict(_C.last_stats())
