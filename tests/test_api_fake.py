"""API behaviour of torchdistx_b200.fake -- the cases the reference pins in
tests/python/test_fake.py:13-60, imported through the drop-in `torchdistx` name, plus geometry
and nesting checks."""
import pytest
import torch

from torchdistx.fake import fake_mode, is_fake, meta_like

no_cuda = pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without CUDA")


@no_cuda
def test_fake_cuda_tensor_reports_cuda_device():
    with fake_mode(fake_cuda=True):
        t = torch.ones([10], device="cuda")
    assert t.device.type == "cuda"
    assert is_fake(t)


@no_cuda
def test_cuda_factory_fails_without_fake_cuda():
    with pytest.raises((AssertionError, RuntimeError)):
        with fake_mode():
            torch.ones([10], device="cuda")


@no_cuda
def test_cuda_factory_fails_again_after_leaving_fake_mode():
    with fake_mode(fake_cuda=True):
        torch.ones([10], device="cuda")
    with pytest.raises((AssertionError, RuntimeError)):
        torch.ones([10], device="cuda")


def test_meta_like_keeps_geometry_and_is_not_fake():
    with fake_mode():
        a = torch.ones([10])
    b = meta_like(a)
    assert not is_fake(b)
    assert b.device.type == "meta"
    assert (b.dtype, b.size(), b.stride()) == (a.dtype, a.size(), a.stride())


def test_meta_like_rejects_real_tensors():
    with pytest.raises(ValueError):
        meta_like(torch.ones([10]))


def test_fake_tensors_have_no_storage_and_infer_shapes():
    with fake_mode():
        a = torch.empty(4, 6, dtype=torch.bfloat16)
        b = (a @ a.t()).float().sum(dim=0)
    assert is_fake(a) and is_fake(b)
    assert b.shape == (4,) and b.dtype == torch.float32
    with pytest.raises((RuntimeError, NotImplementedError)):
        a.data_ptr()
    assert not is_fake(torch.empty(2))  # mode left: real tensors again


def test_fake_mode_nests():
    with fake_mode():
        with fake_mode():
            a = torch.zeros(3)
        b = torch.zeros(3)  # still fake: outer scope is active
    assert is_fake(a) and is_fake(b)
    assert not is_fake(torch.zeros(3))


def test_in_place_ops_keep_tensor_identity_and_refresh_geometry():
    with fake_mode():
        a = torch.zeros(2, 3)
        b = a.add_(1)
        a.resize_(6)
    assert b is a and a.shape == (6,)


def test_repr_of_fake_tensor():
    with fake_mode():
        a = torch.ones(2, 3, dtype=torch.float16, requires_grad=True)
    assert repr(a) == "tensor(..., size=(2, 3), dtype=torch.float16, requires_grad=True, fake=True)"


def test_ops_without_meta_kernel_raise_not_implemented():
    with fake_mode():
        a = torch.ones(4)
        with pytest.raises(NotImplementedError):
            torch.masked_select(a, a > 0)
