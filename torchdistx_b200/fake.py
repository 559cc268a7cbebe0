"""Fake tensors: tensors with a real device, dtype and shape but no storage.

Drop-in for ``torchdistx.fake`` (reference src/python/torchdistx/fake.py:43-82): same three
public names, same signatures, same errors.  Importing this module also teaches
``Tensor.__repr__`` to print fake tensors as ``tensor(..., size=(...), fake=True)``
(reference fake.py:17-40) -- a fake tensor has no data to print.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Iterator

import torch

from . import _C

__all__ = ["fake_mode", "is_fake", "meta_like"]


def _install_repr() -> None:
    if getattr(torch.Tensor.__repr__, "_tdx_fake_aware", False):
        return
    original = torch.Tensor.__repr__

    def __repr__(self: torch.Tensor, *args, **kwargs) -> str:
        if not _C.is_fake(self):
            return original(self, *args, **kwargs)
        fields = [f"size={tuple(self.shape)}"]
        if self.dtype != torch.get_default_dtype():
            fields.append(f"dtype={self.dtype}")
        if self.device.type != "cpu":
            fields.append(f"device={self.device}")
        if self.requires_grad:
            fields.append("requires_grad=True")
        return "tensor(..., " + ", ".join(fields) + ", fake=True)"

    __repr__._tdx_fake_aware = True  # type: ignore[attr-defined]
    torch.Tensor.__repr__ = __repr__  # type: ignore[method-assign]


_install_repr()


@contextmanager
def fake_mode(*, fake_cuda: bool = False) -> Iterator[None]:
    """Every tensor constructed inside the ``with`` block is fake.

    Args:
        fake_cuda: allow ``device="cuda"`` even on a machine without CUDA (ignored when CUDA is
            available).
    """
    _C.enter_fake_mode(fake_cuda)
    try:
        yield
    finally:
        _C.leave_fake_mode()


def is_fake(tensor: torch.Tensor) -> bool:
    """``True`` if ``tensor`` is a fake tensor."""
    return _C.is_fake(tensor)


def meta_like(fake: torch.Tensor) -> torch.Tensor:
    """A meta tensor with the geometry of ``fake``, detached from autograd (like ``detach()``)."""
    try:
        return _C.meta_like(fake)
    except ValueError:
        raise ValueError("`fake` was expected to be a fake tensor.") from None
