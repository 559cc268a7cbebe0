"""InitPlan on the GPU: a saved plan, replayed through the C ABI alone, reproduces
materialize_module bit for bit -- unsharded and sharded."""
import pytest
import torch

from oracle import cases
from torchdistx_b200.deferred_init import deferred_init, materialize_module
from torchdistx_b200.plan import InitPlan

pytestmark = pytest.mark.gpu


def build(case, dtype):
    torch.set_default_dtype(cases.DTYPES[dtype])
    try:
        return deferred_init(lambda: cases.build(case, dtype, "cuda"))
    finally:
        torch.set_default_dtype(torch.float32)


@pytest.mark.parametrize("case,dtype", [("tiny_llama", "bf16"), ("init_zoo", "fp32"), ("tiny_gpt2", "fp32")])
@pytest.mark.parametrize("shard", [None, (1, 4)])
def test_saved_plan_reproduces_materialize_module(case, dtype, shard, tmp_path):
    plan = InitPlan.from_module(build(case, dtype))
    plan.save(str(tmp_path / "p.json"))
    torch.manual_seed(17)
    got = InitPlan.load(str(tmp_path / "p.json")).materialize(device="cuda", shard=shard)
    off_plan = torch.cuda.default_generators[0].get_offset()

    m = build(case, dtype)
    torch.manual_seed(17)
    materialize_module(m, shard=shard)
    off_eng = torch.cuda.default_generators[0].get_offset()
    ref = dict(list(m.named_parameters()) + list(m.named_buffers()))
    assert off_plan == off_eng  # both consumed the generator identically
    for name, t in ref.items():
        if not t.is_cuda:
            continue  # torch.tensor(0)-style real tensors are not part of a plan's GPU output
        g = got[name]
        assert g.shape == t.shape and g.dtype == t.dtype, name
        assert torch.equal(g.detach().reshape(-1).view(torch.uint8), t.detach().contiguous().reshape(-1).view(torch.uint8)), name
        assert isinstance(g, torch.nn.Parameter) == isinstance(t, torch.nn.Parameter)
    if case == "tiny_gpt2":
        assert got["lm_head.weight"] is got["transformer.wte.weight"]


def test_plan_applies_to_a_meta_module():
    plan = InitPlan.from_module(build("mlp_stack", "fp32"))
    with torch.device("meta"):
        skeleton = cases.build("mlp_stack", "fp32")
    torch.manual_seed(3)
    plan.apply(skeleton, device="cuda")
    assert all(p.is_cuda for p in skeleton.parameters())
    y = skeleton(torch.randn(2, 64, device="cuda"))
    assert torch.isfinite(y).all()
