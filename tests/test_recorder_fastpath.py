"""the recorder's per-op overhead.  `normal_` / `uniform_` on a fake tensor return
`self` without a trip through the Meta backend (torch 2.11's Meta kernels for the two are Python:
~200 us and ~14 us per call; the reference pays them for every recorded RNG op,
reference src/cc/torchdistx/fake.cc:476-489).  Semantics must not change."""
import time

import pytest
import torch
from torch import nn

from torchdistx_b200.deferred_init import deferred_init, materialize_module, plan_report
from torchdistx_b200.fake import fake_mode, is_fake


def test_inplace_rng_on_fake_tensors_keeps_eager_semantics():
    def build():
        m = nn.Module()
        w = torch.empty(8, 4)
        assert w.normal_(0.0, 0.5) is w and w.uniform_(-1, 1) is w and is_fake(w)
        assert w.shape == (8, 4) and w.dtype == torch.float32 and w._version == 2
        m.w = nn.Parameter(w)
        m.h = nn.Parameter(torch.empty(3, dtype=torch.float16).normal_(1.0, 0))
        return m

    m = deferred_init(build)
    r = plan_report(m)
    assert r["w"]["source"] == "uniform" and r["w"]["rng_ops"] == 2 and (r["w"]["p0"], r["w"]["p1"]) == (-1.0, 1.0)
    torch.manual_seed(0)
    materialize_module(m)
    torch.manual_seed(0)
    e = torch.empty(8, 4)
    e.normal_(0.0, 0.5)
    e.uniform_(-1, 1)
    assert torch.equal(m.w.detach(), e) and torch.equal(m.h.detach(), torch.ones(3, dtype=torch.float16))


def test_argument_errors_are_raised_while_recording_like_eager():
    with fake_mode():
        t = torch.empty(4)
        with pytest.raises(RuntimeError, match="normal expects std >= 0.0"):
            t.normal_(0.0, -1.0)
        with pytest.raises(RuntimeError, match=r"uniform_ expects to return a \[from, to\) range"):
            t.uniform_(1.0, 0.0)
        i = torch.empty(4, dtype=torch.int64)
        with pytest.raises(Exception):  # integer tensors take the Meta kernel and fail like eager does
            i.normal_(0.0, 1.0)
            torch.empty(4, dtype=torch.int64, device="cpu")  # (not reached)
    eager = torch.empty(4, dtype=torch.int64)
    with pytest.raises(Exception):
        eager.normal_(0.0, 1.0)


def test_recording_an_rng_op_costs_less_than_its_meta_kernel():
    """Relative, so that it holds on any host: 300 recorded `normal_` calls (handler + tape append)
    against 300 plain Meta-device calls (what shape inference alone costs the reference per op)."""
    n = 300

    def record():
        ts = [torch.empty(64, 64) for _ in range(n)]
        t0 = time.perf_counter()
        for t in ts:
            t.normal_(0.0, 0.02)
        record.dt = time.perf_counter() - t0
        m = nn.Module()
        m.p = nn.Parameter(ts[-1])
        return m

    deferred_init(record)
    deferred_init(record)
    ms = [torch.empty(64, 64, device="meta") for _ in range(n)]
    ms[0].normal_(0.0, 0.02)
    t0 = time.perf_counter()
    for t in ms:
        t.normal_(0.0, 0.02)
    meta_dt = time.perf_counter() - t0
    assert record.dt < 0.5 * meta_dt, (record.dt / n * 1e6, meta_dt / n * 1e6)
