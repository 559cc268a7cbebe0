"""InitPlan without a GPU: building, embedding of unfusable tensors, JSON round trip."""
import json

import torch

from oracle import cases
from torchdistx_b200.deferred_init import deferred_init
from torchdistx_b200.plan import FORMAT, InitPlan, offset_increment


def test_plan_of_tiny_llama_round_trips_through_json(tmp_path):
    m = deferred_init(lambda: cases.build("tiny_llama", "bf16"))
    plan = InitPlan.from_module(m)
    by_name = {e.name: e for e in plan.entries}
    w = by_name["model.layers.0.mlp.up_proj.weight"]
    assert (w.source, w.dtype, w.p0, w.p1, w.rng_numels) == ("normal", "BFloat16", 0.0, 0.02, [128 * 64] * 2)
    assert by_name["model.norm.weight"].source == "const"
    inv = by_name["model.rotary_emb.inv_freq"]  # arange -> float -> / dim -> base ** x -> 1 / x -> * 1.0: an index program
    assert inv.source == "iota" and inv.kind == "buffer" and (inv.p0, inv.p1) == (0.0, 2.0)
    assert [e[0] for e in inv.epilogue] == [1, 5, 6, 1] and inv.epilogue[1][1] == 10000.0
    assert plan.num_params == sum(p.numel() for p in m.parameters())
    path = tmp_path / "plan.json"
    plan.save(str(path))
    assert json.load(open(path))["format"] == FORMAT
    again = InitPlan.load(str(path))
    assert [e.__dict__ for e in again.entries] == [e.__dict__ for e in plan.entries]


def test_plan_records_tied_parameters_as_aliases():
    m = deferred_init(lambda: cases.build("tiny_gpt2", "fp32"))
    names = [e.name for e in InitPlan.from_module(m).entries]
    assert names.index("transformer.wte.weight") < names.index("lm_head.weight")  # children first
    e = {e.name: e for e in InitPlan.from_module(deferred_init(lambda: cases.build("tiny_gpt2", "fp32"))).entries}
    assert e["lm_head.weight"].source == "alias" and e["lm_head.weight"].alias_of == "transformer.wte.weight"


def test_offset_increment_matches_the_oracle():
    from oracle import tdx_oracle as O

    for n in (1, 3, 4, 5, 16, 17, 4096, 10**9 + 7):
        assert offset_increment(n) == O.offset_increment(n) and offset_increment(n) % 4 == 0
