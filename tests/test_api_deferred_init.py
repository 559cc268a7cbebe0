"""API behaviour of torchdistx_b200.deferred_init on the CPU (generic replay): the cases the
reference pins in tests/python/test_deferred_init.py:21-75 plus BASELINE config #1
(`deferred_init(nn.Linear(128,128))` == eager same-seed init, bit-exact) and the recording
semantics SURVEY.md section 3.4 lists."""
import pytest
import torch
from torch import nn
from torch.nn import Module, Parameter

from torchdistx.deferred_init import deferred_init, is_deferred, materialize_module, materialize_tensor
from torchdistx.fake import is_fake
from torchdistx_b200.deferred_init import last_materialize_stats


class TwoParams(Module):
    def __init__(self):
        super().__init__()
        self.param1 = Parameter(torch.ones([5]))
        self.param2 = Parameter(torch.ones([5]))


class Tied(Module):
    def __init__(self):
        super().__init__()
        self.param1 = Parameter(torch.ones([5]))
        self.param2 = self.param1


def test_real_tensor_is_returned_unchanged():
    a = torch.ones([10])
    assert materialize_tensor(a) is a


def test_repeated_and_aliased_materialize_return_one_object():
    module = deferred_init(Tied)
    a = materialize_tensor(module.param1)
    b = materialize_tensor(module.param1)
    c = materialize_tensor(module.param2)
    assert a is b and a is c
    assert isinstance(a, Parameter) and not is_fake(a)
    assert torch.equal(a, torch.ones(5))


def test_is_deferred_transitions():
    assert not is_deferred(TwoParams())
    module = deferred_init(TwoParams)
    assert is_deferred(module)
    materialize_module(module)
    assert not is_deferred(module)

    module = deferred_init(TwoParams)
    module.param1 = materialize_tensor(module.param1)
    assert is_deferred(module)
    module.param2 = materialize_tensor(module.param2)
    assert not is_deferred(module)


def test_is_deferred_rejects_other_types():
    with pytest.raises(ValueError):
        is_deferred(3)


def test_materialize_tensor_rejects_non_tensors():
    with pytest.raises(TypeError):
        materialize_tensor("weight")


def test_linear_128_equals_eager_same_seed_bit_exact():
    # BASELINE.json configs[0]
    torch.manual_seed(0)
    m = deferred_init(nn.Linear, 128, 128)
    assert is_fake(m.weight) and m.weight.shape == (128, 128)
    torch.manual_seed(0)
    materialize_module(m)
    torch.manual_seed(0)
    e = nn.Linear(128, 128)
    assert torch.equal(m.weight, e.weight) and torch.equal(m.bias, e.bias)
    assert isinstance(m.weight, Parameter) and m.weight.requires_grad
    assert last_materialize_stats()["generic_ops"] > 0  # CPU tensors replay through ATen
    m(torch.ones(2, 128)).sum().backward()
    assert m.weight.grad is not None


def test_rng_follows_materialize_order_and_dead_ops_consume():
    # bias-then-weight differs from eager; a dead uniform_ still advances mt19937
    def build():
        lin = nn.Linear(16, 16, bias=False)
        nn.init.normal_(lin.weight, 0.0, 0.02)
        return lin

    m = deferred_init(build)
    torch.manual_seed(7)
    materialize_module(m)
    torch.manual_seed(7)
    e = build()
    assert torch.equal(m.weight, e.weight)


def test_buffers_only_and_check_fn():
    class Net(Module):
        def __init__(self):
            super().__init__()
            self.bn = nn.BatchNorm1d(4)
            self.fc = nn.Linear(4, 4)

    m = deferred_init(Net)
    materialize_module(m, buffers_only=True)
    assert not is_fake(m.bn.running_mean) and is_fake(m.bn.weight) and is_fake(m.fc.weight)
    materialize_module(m, check_fn=lambda mod: not isinstance(mod, nn.Linear))
    assert not is_fake(m.bn.weight) and is_fake(m.fc.weight)
    materialize_module(m)
    assert not is_deferred(m)
    assert torch.equal(m.bn.running_var, torch.ones(4))


def test_views_and_in_place_ops_through_views():
    class V(Module):
        def __init__(self):
            super().__init__()
            w = torch.zeros(4, 4)
            w.view(-1)[::5] = 1.0  # identity through a strided view
            w[0].mul_(3.0)
            self.w = Parameter(w)

    m = deferred_init(V)
    materialize_module(m)
    expect = torch.eye(4)
    expect[0] *= 3
    assert torch.equal(m.w, expect)


def test_external_tensor_version_check():
    ext = torch.ones(3)

    def build():
        return Parameter(torch.zeros(3) + ext)

    p = deferred_init(build)
    ext.add_(1)
    with pytest.raises(RuntimeError, match="updated in-place"):
        materialize_tensor(p)


def test_item_is_terminal_and_keeps_recording():
    def build():
        a = torch.full((3,), 2.0)
        s = a.sum().item()  # needs a real value now
        return Parameter(a * s)

    p = deferred_init(build)
    assert is_fake(p)
    assert torch.equal(materialize_tensor(p), torch.full((3,), 12.0))


def test_a_dead_draw_that_reads_a_constant_tensor_keeps_its_place_in_the_stream():
    """`torch.randn(n).copy_(a)`: the randn is dead, but it is drawn -- when `a` is replayed, because
    the copy_ reads `a` and runs with a's history (as in the reference, whose call stack takes it in
    as a dependent of a's storage).  `a` itself draws nothing, and RNG-free programs are otherwise
    replayed at the END of a materialize_module call: that moved this draw behind `b`'s."""
    def build():
        m = Module()
        a = torch.full((4, 4), 1.0)
        dead = torch.randn(4, 4).copy_(a)
        b = torch.empty(4, 4).uniform_()
        m.a, m.b, m.dead = Parameter(a), Parameter(b), Parameter(dead)
        return m

    m = deferred_init(build)
    torch.manual_seed(3)
    materialize_module(m)
    torch.manual_seed(3)
    e = build()
    assert torch.equal(m.a, e.a) and torch.equal(m.dead, e.dead) and torch.equal(m.b, e.b)


def test_a_failing_constructor_leaves_the_thread_out_of_deferred_mode():
    def boom():
        nn.Linear(4, 4)
        raise KeyError("constructor failed")

    with pytest.raises(KeyError):
        deferred_init(boom)
    assert not is_fake(torch.ones(2))  # (reference deferred_init.py:36-41: leave in a `finally`)

    def outer():  # a failure in a NESTED deferred_init leaves the outer recording running
        with pytest.raises(KeyError):
            deferred_init(boom)
        return nn.Linear(3, 3)

    m = deferred_init(outer)
    assert is_deferred(m) and not is_fake(torch.ones(1))
    torch.manual_seed(0)
    materialize_module(m)
    torch.manual_seed(0)
    assert torch.equal(m.weight, nn.Linear(3, 3).weight)


def test_tolist_and_numpy_read_a_deferred_tensor_like_item_does():
    """Constructors that turn a tensor into Python numbers (stochastic-depth rates:
    `torch.linspace(0, rate, depth).tolist()`): the value is needed now, so the tensor is built now
    and recording goes on.  (The reference supports `item()` only: `tolist()` / `numpy()` read memory
    without passing the dispatcher and fail on a tensor without storage.)"""
    def build():
        m = nn.Linear(4, 4)
        rates = torch.linspace(0, 0.3, 4).tolist()
        scale = torch.full((2,), 3.0).numpy()
        m.rates, m.scale = rates, float(scale[0])
        m.extra = Parameter(torch.empty(4, 4).normal_() * m.scale + rates[1])
        return m

    m = deferred_init(build)
    assert m.rates == torch.linspace(0, 0.3, 4).tolist() and m.scale == 3.0 and is_deferred(m)
    torch.manual_seed(0)
    materialize_module(m)
    torch.manual_seed(0)
    e = build()
    assert torch.equal(m.weight, e.weight) and torch.equal(m.bias, e.bias) and torch.equal(m.extra, e.extra)
    # real tensors are untouched by the wrappers; a plain fake tensor (no recording) still has no data to give
    assert torch.ones(2).tolist() == [1.0, 1.0] and torch.ones(2).numpy().tolist() == [1.0, 1.0]
    from torchdistx_b200.fake import fake_mode
    with fake_mode():
        f = torch.ones(2)
    with pytest.raises(RuntimeError):
        f.tolist()


def test_data_setter_and_getter_are_recorded():
    def build():
        lin = nn.Linear(3, 3)
        lin.weight.data.fill_(0.5)
        lin.bias.data = torch.arange(3.0)
        return lin

    m = deferred_init(build)
    materialize_module(m)
    assert torch.equal(m.weight, torch.full((3, 3), 0.5))
    assert torch.equal(m.bias, torch.arange(3.0))


def test_default_dtype_and_explicit_generator_are_honoured():
    g = torch.Generator().manual_seed(123)

    def build():
        return Parameter(torch.empty(64).normal_(generator=g))

    p = deferred_init(build)
    out = materialize_tensor(p)
    g2 = torch.Generator().manual_seed(123)
    assert torch.equal(out, torch.empty(64).normal_(generator=g2))


def test_fake_tensor_from_plain_fake_mode_is_rejected():
    from torchdistx.fake import fake_mode

    with fake_mode():
        a = torch.ones(3)
    with pytest.raises(ValueError):
        deferred_init(lambda: a + 1)


def test_nested_deferred_init_and_cross_scope_inputs():
    emb = deferred_init(nn.Embedding, 10, 4)

    def build():
        return Parameter(emb.weight.detach() * 2.0)

    p = deferred_init(build)
    torch.manual_seed(3)
    out = materialize_tensor(p)
    w = materialize_tensor(emb.weight)
    assert torch.equal(out, w.detach() * 2.0)


def test_cross_scope_input_without_random_draws_is_built_on_demand():
    """A tensor of a LATER deferred_init that reads one of an EARLIER one, materialised first: the
    earlier tensor is built on the spot as its argument -- also when its program draws no random
    numbers (such programs are otherwise replayed at the end of a call; as an argument that handed
    back nothing: "Expected a proper Tensor but got None")."""
    for src_fn in (lambda: torch.full((4, 3), 2.0), lambda: Parameter(torch.ones(4, 3) * 2.0),
                   lambda: torch.tril(torch.ones(4, 3)) * 2.0):
        for reader in (lambda a: a * 2.0 + 1.0, lambda a: a.clone(), lambda a: a.detach().clone().mul_(3.0),
                       lambda a: torch.zeros(4, 3).copy_(a)):
            a = deferred_init(src_fn)
            b = deferred_init(lambda: Parameter(reader(a.detach())))
            out = materialize_tensor(b)  # (a is still fake here)
            assert torch.equal(out.detach(), reader(src_fn().detach()))
            assert torch.equal(materialize_tensor(a).detach(), src_fn().detach())


def test_materialize_module_error_leaves_the_module_untouched_and_does_not_hang():
    """materialize_module plans on a helper thread; a failure there must surface as the same
    exception on the calling thread, with no tensor assigned and the next call working."""
    ext = torch.ones(4)

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = nn.Parameter(torch.zeros(3))
            self.b = nn.Parameter(torch.full((4,), 2.0) + ext)   # replay needs `ext` unchanged
            self.c = nn.Parameter(torch.ones(2))

    m = deferred_init(M)
    ext.add_(1)
    with pytest.raises(RuntimeError, match="updated in-place"):
        materialize_module(m)
    assert is_fake(m.a) and is_fake(m.b) and is_fake(m.c)  # nothing was assigned
    ok = deferred_init(nn.Linear, 3, 2)
    materialize_module(ok)
    assert not is_deferred(ok)


def test_materialize_module_runs_under_the_callers_thread_local_state():
    """The helper thread applies the caller's ThreadLocalState: results built under no_grad /
    with another default dtype behave as if built inline."""
    m = deferred_init(nn.Linear, 4, 4)
    with torch.no_grad():
        materialize_module(m)
    assert m.weight.requires_grad and m.weight.grad_fn is None
    torch.manual_seed(3)
    a = deferred_init(nn.Linear, 8, 8)
    materialize_module(a)
    torch.manual_seed(3)
    b = nn.Linear(8, 8)
    assert torch.equal(a.weight, b.weight) and torch.equal(a.bias, b.bias)  # generator written back before return


def test_materialize_module_in_a_forked_child():
    """The helper thread does not survive fork(); the child must start its own."""
    import os
    materialize_module(deferred_init(nn.Linear, 2, 2))  # make sure the parent's helper exists
    pid = os.fork()
    if pid == 0:
        try:
            m = deferred_init(nn.Linear, 3, 3)
            materialize_module(m)
            os._exit(0 if not is_deferred(m) else 3)
        except BaseException:
            os._exit(4)
    _, status = os.waitpid(pid, 0)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0


def test_materialize_module_from_two_python_threads():
    """Two Python threads share the one helper thread: calls serialise there, neither deadlocks, and
    each thread's recording (deferred_init state is thread-local) comes back complete."""
    import threading

    errors, done = [], []

    def work(seed):
        try:
            for i in range(10):
                m = deferred_init(lambda: nn.Sequential(nn.Linear(16, 16), nn.LayerNorm(16), nn.Linear(16, 4)))
                materialize_module(m)
                assert not is_deferred(m)
                assert torch.equal(m[1].weight, torch.ones(16)) and m[0].weight.shape == (16, 16)
            done.append(seed)
        except BaseException as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=work, args=(s,)) for s in (1, 2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errors, errors
    assert sorted(done) == [1, 2]
