"""Random sequences of the public calls on a module tree with tied parameters, buffers computed
from parameters, `None` entries and writes through views: `materialize_tensor` on random tensors
(twice: same object), `materialize_module` on random submodules with `buffers_only` / `check_fn`, then
the whole tree.  Whatever the sequence: nothing is left deferred, every value is the eager one (the
programs draw no random numbers, so order cannot matter), parameters stay `Parameter`s with their
`requires_grad`, tied parameters stay ONE object under all their names, and a tensor handed out by
`materialize_tensor` earlier is the very object the module ends up holding (reference
tests/python/test_deferred_init.py:21-44, `_C/deferred_init.cc:80-94`)."""
import random

import torch
from torch import nn

from torchdistx_b200.deferred_init import deferred_init, is_deferred, materialize_module, materialize_tensor
from torchdistx_b200.fake import is_fake


class Leaf(nn.Module):
    def __init__(self, r, shared):
        super().__init__()
        d = r.choice([4, 6])
        self.w = nn.Parameter(torch.full((d, 3), r.choice([0.5, 2.0, -1.0])))
        with torch.no_grad():
            if r.random() < 0.5: self.w[1].zero_()
            if r.random() < 0.3: self.w.mul_(3.0)
        if r.random() < 0.5: self.register_buffer("b", torch.arange(d).float() * 0.5)
        if r.random() < 0.3: self.register_buffer("c", self.w.detach() * 2 + 1)   # buffer computed from the parameter
        if shared is not None and r.random() < 0.4: self.tied = shared                   # tied parameter
        if r.random() < 0.2: self.none_param = None

class Tree(nn.Module):
    def __init__(self, seed):
        super().__init__()
        r = random.Random(seed)
        self.shared = nn.Parameter(torch.ones(2, 2) * 7)
        self.a = Leaf(r, self.shared)
        self.mid = nn.Sequential(Leaf(r, self.shared), nn.ModuleList([Leaf(r, None), Leaf(r, self.shared)]))
        self.z = Leaf(r, None)
        self.register_buffer("top", torch.zeros(3))

def one(seed):
    r = random.Random(seed * 7 + 1)
    m = deferred_init(Tree, seed)
    e = Tree(seed)
    mods = dict(m.named_modules())
    handed = {}
    for _ in range(r.randint(1, 6)):
        act = r.choice(["tensor", "module", "module_buf", "module_check", "again"])
        if act == "tensor":
            name, t = r.choice(list(m.named_parameters()) + list(m.named_buffers()))
            out = materialize_tensor(t)
            if name in handed: assert out is handed[name], ("identity", name)
            handed[name] = out
            assert not is_fake(out)
        elif act == "again" and handed:
            name = r.choice(list(handed)); owner, _, key = name.rpartition(".")
            # the module still holds the fake (materialize_tensor does not assign); re-materialising gives the same object
            mod = mods[owner]
            t = mod._parameters.get(key, None) if key in mod._parameters else mod._buffers.get(key)
            if t is not None:
                assert materialize_tensor(t) is handed[name] or not is_fake(t), ("identity2", name)
        else:
            target = r.choice(list(mods.values()))
            kw = {}
            if act == "module_buf": kw["buffers_only"] = True
            if act == "module_check":
                skip = r.choice(list(mods.values()))
                kw["check_fn"] = lambda mod, skip=skip: mod is not skip
            materialize_module(target, **kw)
    materialize_module(m)
    assert not is_deferred(m)
    # values, classes, tying
    me, ee = dict(m.named_parameters(remove_duplicate=False)), dict(e.named_parameters(remove_duplicate=False))
    assert set(me) == set(ee)
    for k in ee:
        assert isinstance(me[k], nn.Parameter) and torch.equal(me[k].detach(), ee[k].detach()), ("value", k)
        assert me[k].requires_grad == ee[k].requires_grad
    mb, eb = dict(m.named_buffers()), dict(e.named_buffers())
    assert set(mb) == set(eb)
    for k in eb: assert torch.equal(mb[k], eb[k]), ("buffer", k)
    ties_m = {k for k in me if me[k] is m.shared}; ties_e = {k for k in ee if ee[k] is e.shared}
    assert ties_m == ties_e, ("tying", ties_m, ties_e)
    for name, out in handed.items():
        cur = me.get(name, mb.get(name))
        assert cur is out, ("handed identity", name)


def test_any_sequence_of_public_calls_ends_in_the_eager_module():
    for seed in range(600):
        try:
            one(seed)
        except AssertionError as e:
            raise AssertionError(f"seed {seed}: {e}") from e
