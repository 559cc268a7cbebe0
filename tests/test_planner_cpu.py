"""Planner verdicts without a GPU (`plan_report` evaluates the recorded programs symbolically and
allocates nothing): the init idioms of SURVEY.md section 3.4 fold into single descriptors, dead RNG
passes are seen, and whole model families stay on the fused path."""
import math

import pytest
import torch
from torch import nn

from oracle import cases
from torchdistx_b200.deferred_init import deferred_init, plan_report


def report(fn):
    return plan_report(deferred_init(fn))


def test_init_idioms_fold_as_documented():
    r = report(lambda: cases.build("init_zoo", "fp32"))
    k = r["kaiming.weight"]
    assert k["source"] == "uniform" and k["rng_ops"] == 1 and k["n_epilogue"] == 0
    bound = 1 / math.sqrt(48)
    assert k["p0"] == pytest.approx(-bound, rel=1e-6) and k["p1"] == pytest.approx(bound, rel=1e-6)
    assert r["embed.weight"]["source"] == "normal" and (r["embed.weight"]["p0"], r["embed.weight"]["p1"]) == (0.0, 1.0)
    assert r["norm.weight"]["source"] == "const" and r["norm.bias"]["source"] == "const"
    t = r["trunc"]  # uniform_ -> erfinv_ -> mul_ -> add_ -> clamp_
    assert t["source"] == "uniform" and t["n_epilogue"] == 4 and t["fusible"]
    tw = r["twice.weight"]  # Linear's kaiming uniform_ overwritten by normal_: one dead pass
    assert tw["source"] == "normal" and tw["rng_ops"] == 2 and tw["p1"] == pytest.approx(0.02)
    assert r["const"]["source"] == "const"  # ones * 3 + 1 folded through ATen
    s = r["scaled"]  # randn * 0.02 + 1.0
    assert s["source"] == "normal" and s["n_epilogue"] == 2
    assert r["int_fill"]["source"] == "const" and r["int_fill"]["dtype"] == "Long"
    assert r["mask"]["source"] == "opaque" and not r["mask"]["fusible"]  # tril(...).bool(): replayed by ATen on the target device
    assert r["steps"]["source"] == "iota" and r["steps"]["fusible"] and (r["steps"]["p0"], r["steps"]["p1"]) == (0.0, 1.0)
    assert r["bn.num_batches_tracked"]["source"] == "real"  # torch.tensor(0) is never intercepted


def test_rotary_buffers_of_a_converted_model_stay_index_programs():
    """`LlamaForCausalLM(...).to(torch.bfloat16)` converts the fp32 rotary buffers too: arange -> ... ->
    reciprocal -> to(bf16) is still one iota descriptor (fp32 program, one rounding at the store)."""
    import torch

    r = report(lambda: cases.build("tiny_llama", "fp32").to(torch.bfloat16))
    for name in ("model.rotary_emb.inv_freq", "model.rotary_emb.original_inv_freq"):
        assert (r[name]["source"], r[name]["dtype"], r[name]["fusible"]) == ("iota", "BFloat16", True), r[name]
    assert fusible_fraction(r) == 1.0


def test_llama_linear_chain_has_one_dead_uniform():
    r = report(lambda: cases.build("tiny_llama", "bf16"))
    w = r["model.layers.0.mlp.up_proj.weight"]
    assert (w["source"], w["dtype"], w["rng_ops"], w["p0"], w["p1"]) == ("normal", "BFloat16", 2, 0.0, pytest.approx(0.02))
    assert r["model.norm.weight"]["source"] == "const"


def fusible_fraction(r):
    total = sum(v["numel"] for v in r.values())
    return sum(v["numel"] for v in r.values() if v["fusible"]) / total


@pytest.mark.parametrize("case", ["tiny_llama", "tiny_gpt2", "torch_transformer", "mlp_stack", "clones"])
def test_model_families_stay_on_the_fused_path(case):
    assert fusible_fraction(report(lambda: cases.build(case, "fp32"))) > 0.999


def test_more_hf_families_and_torch_modules():
    import transformers as T

    fams = {
        "vit": lambda: T.ViTModel(T.ViTConfig(hidden_size=64, num_hidden_layers=1, num_attention_heads=4,
                                              intermediate_size=128, image_size=32, patch_size=8)),
        "t5": lambda: T.T5Model(T.T5Config(d_model=64, d_ff=128, num_layers=1, num_heads=4, vocab_size=512)),
        "mixtral": lambda: T.MixtralForCausalLM(T.MixtralConfig(vocab_size=512, hidden_size=64, intermediate_size=128,
                                                                num_hidden_layers=1, num_attention_heads=4,
                                                                num_key_value_heads=2, num_local_experts=2)),
        "lstm": lambda: nn.LSTM(32, 64, num_layers=2),
        "conv_bn": lambda: nn.Sequential(nn.Conv2d(3, 16, 3), nn.BatchNorm2d(16), nn.Linear(10, 10)),
    }
    for name, fn in fams.items():
        assert fusible_fraction(report(fn)) > 0.99, name


def test_mid_history_reader_forces_generic_replay():
    def build(reader):
        w = torch.empty(8, 8).uniform_()
        snapshot = reader(w)  # reads the uniform state ...
        w.normal_()  # ... which is then overwritten
        m = nn.Module()
        m.w, m.s = nn.Parameter(w), nn.Parameter(snapshot)
        return m

    from torchdistx_b200.deferred_init import materialize_module

    # a reader the planner cannot express must run at its own point in history: w is replayed
    m = deferred_init(build, torch.sin)
    r = plan_report(m)
    assert r["w"]["source"] == "opaque" and r["s"]["source"] == "opaque"
    torch.manual_seed(0)
    materialize_module(m)
    torch.manual_seed(0)
    e = build(torch.sin)
    assert torch.equal(m.w, e.w) and torch.equal(m.s, e.s)

    # a reader whose result folds symbolically (uniform * 2: the state of w AS OF the reader) pins nothing
    m = deferred_init(build, lambda w: w * 2.0)
    r = plan_report(m)
    assert (r["w"]["source"], r["w"]["rng_ops"]) == ("normal", 2)
    assert (r["s"]["source"], r["s"]["n_epilogue"], r["s"]["rng_ops"]) == ("uniform", 1, 1)
    torch.manual_seed(0)
    materialize_module(m)  # (CPU tensors: ATen replay either way, bit-exact with eager)
    torch.manual_seed(0)
    e = build(lambda w: w * 2.0)
    assert torch.equal(m.w, e.w) and torch.equal(m.s, e.s)


def test_to_dtype_variant_is_fused_with_the_fp32_stream():
    r = report(lambda: cases.build("cast_variant", "fp32"))
    assert fusible_fraction(r) == 1.0
    # Module.to(bf16) rebinds the parameter to its cast and nothing names the fp32 tensor any more:
    # "equal to fp32_tensor.to(bf16)" is unobservable, the tensor takes the native 16-bit stream
    w = r["body.0.weight"]  # Linear(fp32) -> Module.to(bf16): uniform_ -> _to_copy -> set_data
    assert (w["source"], w["dtype"], w["wide"]) == ("uniform", "BFloat16", False)
    h = r["head.weight"]  # kaiming (dead) -> trunc_normal_ chain -> half()
    assert (h["source"], h["dtype"], h["wide"], h["n_epilogue"], h["rng_ops"]) == ("uniform", "Half", False, 4, 2)
    # (`a` is still a parameter: `b = a.to(bf16)` must equal it after rounding, bit for bit)
    assert r["body.1.weight"]["source"] == "const" and r["body.1.weight"]["dtype"] == "BFloat16"
    assert r["a"]["wide"] is False and r["b"]["wide"] is True and r["b"]["n_epilogue"] == 1


def test_partial_writes_through_views_become_segments():
    """`padding_idx` embeddings (normal_ then one row zeroed through `weight[i]`) and slice-wise
    initialisation fold into segments of one tensor (reference: the ops are simply replayed in
    order, deferred_init.cc:541-622)."""
    r = report(lambda: cases.build("padded_embeddings", "fp32"))
    e = r["torch_style.weight"]
    assert e["fusible"] and e["source"] == "normal"
    segs = [(s["begin"], s["end"], s["source"]) for s in e["segments"]]
    assert segs == [(0, 3 * 64, "normal"), (3 * 64, 4 * 64, "const"), (4 * 64, 512 * 64, "normal")]
    assert e["segments"][0]["rng_pass"] == e["segments"][2]["rng_pass"] == 0 and e["segments"][2]["origin"] == 0
    h = r["hf_style.weight"]  # Embedding's own normal_ (dead), HF's normal_(0, 0.02) (live), row 0 zeroed
    assert h["fusible"] and h["rng_ops"] == 2 and [s["source"] for s in h["segments"]] == ["const", "normal"]
    assert h["segments"][1]["p1"] == pytest.approx(0.02) and h["segments"][1]["rng_pass"] == 1
    last = r["last_row.weight"]
    assert [s["source"] for s in last["segments"]] == ["normal", "const"]
    hv = r["halves"]  # [:32].normal_(0, .1); [32:].fill_(.5); [:16].mul_(2)
    assert hv["fusible"]
    assert [(s["begin"], s["end"], s["source"], len(s["epilogue"])) for s in hv["segments"]] == [
        (0, 16 * 32, "normal", 1), (16 * 32, 32 * 32, "normal", 0), (32 * 32, 64 * 32, "const", 0)]
    assert fusible_fraction(r) == 1.0


def test_padding_idx_model_families_are_fusible_by_bytes():
    import transformers as T

    fams = {
        "bert": lambda: T.BertModel(T.BertConfig(vocab_size=512, hidden_size=64, num_hidden_layers=1,
                                                 num_attention_heads=4, intermediate_size=128)),
        "opt": lambda: T.OPTForCausalLM(T.OPTConfig(vocab_size=512, hidden_size=64, ffn_dim=128, num_hidden_layers=1,
                                                    num_attention_heads=4, word_embed_proj_dim=64,
                                                    max_position_embeddings=64)),
        "gemma2": lambda: T.Gemma2ForCausalLM(T.Gemma2Config(vocab_size=512, hidden_size=64, intermediate_size=128,
                                                             num_hidden_layers=1, num_attention_heads=4,
                                                             num_key_value_heads=2, head_dim=16)),
        "phi3": lambda: T.Phi3ForCausalLM(T.Phi3Config(vocab_size=512, hidden_size=64, intermediate_size=128,
                                                       num_hidden_layers=1, num_attention_heads=4,
                                                       num_key_value_heads=2, pad_token_id=0)),
    }
    for name, fn in fams.items():
        r = report(fn)
        # every floating-point tensor folds; what is left are `arange`-style integer buffers
        # (BERT's position_ids: 512 int64 of a 100k-element toy config, 4 KB of a real one's 440 MB)
        assert all(v["fusible"] for v in r.values() if v["dtype"] in ("Float", "BFloat16", "Half") and v["numel"] > 64), name
        assert fusible_fraction(r) >= 0.99, name


def test_analysis_is_cached_on_the_recording_and_matches_a_fresh_evaluation():
    from torchdistx_b200 import _C

    m = deferred_init(lambda: cases.build("init_zoo", "fp32"))
    r1 = plan_report(m)
    r2 = plan_report(m)
    assert {k: v["source"] for k, v in r1.items()} == {k: v["source"] for k, v in r2.items()}
    # `ones * 3 + 1` needs the target device's arithmetic: not cached, evaluated on demand, same verdict
    assert r1["const"]["source"] == "const" and r1["const"]["fusible"]
    assert "aten::select [view]" in _C.storage_history(
        deferred_init(lambda: cases.build("padded_embeddings", "fp32")).torch_style.weight)


def test_materialize_module_rewords_only_per_tensor_value_errors():
    m = deferred_init(lambda: cases.build("linear128", "fp32"))
    from torchdistx_b200.deferred_init import materialize_module
    with pytest.raises(ValueError, match=r"^shard must be \(rank, world_size\)[^(]*$"):
        materialize_module(m, shard=(3, 2))
    materialize_module(m)
    assert not plan_report(m)["weight"]["deferred"]


def test_submission_sizes_follow_the_recordings_total():
    """How a materialize_module call cuts its submissions (Batch::note): a first 256 MiB one, then at
    most three more whose sizes grow by the ratio of the host's planning rate to the GPU's writing
    rate -- x4 when one GPU owns the whole model (big tensors: the GPU is the bottleneck), equal
    chunks when a rank owns an eighth (the call ends one chunk's GPU time after the last tensor is
    planned).  A function of byte counts only."""
    import torch  # noqa: F401  (loads the libraries _C links against)

    from torchdistx_b200 import _C

    total, n = 16_060_000_000, 293  # Llama-3-8B bf16
    one = _C._submission_sizes(total, n)
    assert len(one) <= 4 and abs(sum(one) - total) < n
    assert one[0] < 300e6 and one[-1] > 0.6 * total and all(b > 2 * a for a, b in zip(one, one[1:]))
    eighth = _C._submission_sizes(total // 8, n)
    assert len(eighth) == 4 and eighth[0] < 300e6
    assert max(eighth[1:]) < 1.1 * min(eighth[1:]) and eighth[-1] < 0.3 * (total // 8)
    quarter = _C._submission_sizes(total // 4, n)
    assert len(quarter) == 4 and all(1.5 * a < b < 2.5 * a for a, b in zip(quarter[1:], quarter[2:]))
    # no estimate (several recordings in one call): x4 per submission, whatever the sizes
    blind = _C._submission_sizes(total, n, with_estimate=False)
    assert blind[0] < 300e6 and 3.5 < blind[1] / blind[0] < 4.5 and 3.5 < blind[2] / blind[1] < 4.5
    # deterministic
    assert _C._submission_sizes(total // 8, n) == eighth
    # a call below the first threshold is one submission
    assert _C._submission_sizes(50_000_000, 100) == [50_000_000]
