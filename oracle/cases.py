"""Module builders shared by the product tests and the reference oracle (TEST INFRASTRUCTURE).

Each case is `name -> callable() -> nn.Module`, pure PyTorch / transformers, importing neither
torchdistx_b200 nor the reference.  Sizes are small enough for the CPU reference to run in seconds.
"""
from __future__ import annotations

import math

import torch
from torch import nn


def linear128():
    return nn.Linear(128, 128)


class InitZoo(nn.Module):
    """One tensor per init idiom the planner folds."""

    def __init__(self):
        super().__init__()
        self.kaiming = nn.Linear(48, 32)
        self.embed = nn.Embedding(64, 16)
        self.norm = nn.LayerNorm(16)
        self.bn = nn.BatchNorm1d(8)
        self.xavier = nn.Parameter(torch.empty(24, 40))
        nn.init.xavier_uniform_(self.xavier)
        self.xavier_n = nn.Parameter(torch.empty(24, 40))
        nn.init.xavier_normal_(self.xavier_n, gain=0.5)
        self.trunc = nn.Parameter(torch.empty(32, 32))
        nn.init.trunc_normal_(self.trunc, mean=0.1, std=0.02, a=-0.04, b=0.06)
        self.twice = nn.Linear(32, 32, bias=False)
        nn.init.normal_(self.twice.weight, mean=0.0, std=0.02)  # dead uniform_ + live normal_
        self.const = nn.Parameter(torch.ones(17) * 3.0 + 1.0)
        self.scaled = nn.Parameter(torch.randn(33, 7) * 0.02 + 1.0)
        self.zeros = nn.Parameter(torch.zeros(5, 3))
        self.full = nn.Parameter(torch.full((9,), 0.25))
        self.register_buffer("mask", torch.tril(torch.ones(6, 6)).bool())
        self.register_buffer("steps", torch.arange(10))
        self.register_buffer("int_fill", torch.full((4,), 7, dtype=torch.int64))


def init_zoo():
    return InitZoo()


def tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64)
    return LlamaForCausalLM(cfg)


def tiny_gpt2():
    from transformers import GPT2Config, GPT2LMHeadModel

    cfg = GPT2Config(n_layer=2, n_embd=64, n_head=4, vocab_size=512, n_positions=64)
    return GPT2LMHeadModel(cfg)


def torch_transformer():
    # deepcopy'd layers (aten::clone of initialised tensors) + a second init pass over every matrix
    return nn.Transformer(d_model=32, nhead=4, num_encoder_layers=2, num_decoder_layers=1, dim_feedforward=64)


class Clones(nn.Module):
    """deepcopy / .data.clone() of an RNG-initialised parameter must reproduce its values."""

    def __init__(self):
        super().__init__()
        import copy
        self.a = nn.Parameter(torch.empty(40, 24).normal_(0.0, 0.5))
        self.b = copy.deepcopy(self.a)
        self.c = nn.Parameter(self.a.data.clone())
        self.d = nn.Parameter(self.a.detach().clone() * 2.0)


def clones():
    return Clones()


class CastVariant(nn.Module):
    """The `.to(dtype)` idiom: initialised in fp32, converted afterwards (Module.to rebinds
    `param.data`), plus a tensor and its cast both kept alive."""

    def __init__(self):
        super().__init__()
        self.body = nn.Sequential(nn.Linear(48, 40), nn.LayerNorm(40), nn.Embedding(32, 24)).to(torch.bfloat16)
        self.head = nn.Linear(40, 16)
        nn.init.trunc_normal_(self.head.weight, std=0.02, a=-0.05, b=0.05)
        self.head.half()
        self.a = nn.Parameter(torch.randn(64, 33) * 0.02)
        self.b = nn.Parameter(self.a.detach().to(torch.bfloat16))
        self.u = nn.Parameter(torch.empty(50, 20).uniform_(-0.3, 0.1))
        self.v = nn.Parameter(self.u.detach().to(torch.float16))
        # fp32 steps, the cast, then more steps in the 16-bit dtype
        self.w = nn.Parameter(self.a.detach().to(torch.bfloat16).mul_(3.0).add_(1.0))


def cast_variant():
    return CastVariant()


def mlp_stack():
    return nn.Sequential(nn.Linear(64, 256), nn.GELU(), nn.Linear(256, 64), nn.LayerNorm(64))


# ---- BASELINE-sized cases (tests/test_t1_fullsize_gpu.py; SURVEY.md section 8c tier T1) -----------------
LLAMA8B_LAYER = dict(vocab_size=2048, hidden_size=4096, intermediate_size=14336, num_hidden_layers=1,
                     num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192,
                     rope_theta=500000.0)


def llama8b_layer():
    """The bounded sample bench.py times the reference on: one full-width Llama-3-8B decoder layer
    (all seven linears at their real sizes, 4096x14336 included), vocabulary cut to 2048 rows:
    234,893,312 parameters."""
    from transformers import LlamaConfig, LlamaForCausalLM

    return LlamaForCausalLM(LlamaConfig(**LLAMA8B_LAYER))


def llama8b_layer_cast():
    """SURVEY 8d cfg3 variant: built in fp32, converted with `.to(torch.bfloat16)`."""
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        return llama8b_layer().to(torch.bfloat16)
    finally:
        torch.set_default_dtype(prev)


BIG = 4096  # 4096 x 4096 = 2^24 elements per tensor


class BigInits(nn.Module):
    """One 2^24-element tensor per RNG idiom of BASELINE config 5 (+ the two epilogue chains)."""

    def __init__(self):
        super().__init__()
        self.normal = nn.Parameter(torch.empty(BIG, BIG).normal_(0.0, 0.02))
        self.uniform = nn.Parameter(torch.empty(BIG, BIG).uniform_(-0.05, 0.03))
        self.kaiming = nn.Parameter(torch.empty(BIG, BIG))
        nn.init.kaiming_uniform_(self.kaiming, a=math.sqrt(5))  # bound = 1/sqrt(fan_in) = 1/64
        self.trunc = nn.Parameter(torch.empty(BIG, BIG))
        nn.init.trunc_normal_(self.trunc, mean=0.0, std=0.02, a=-0.04, b=0.04)
        self.scaled = nn.Parameter(torch.randn(BIG, BIG) * 0.02 + 1.0)


def big_inits():
    return BigInits()


class PaddedEmbeddings(nn.Module):
    """`padding_idx` embeddings: normal_ then one row zeroed through a view (nn.Embedding itself,
    and the HF `_init_weights` idiom of BERT / OPT / Gemma / Phi-3)."""

    def __init__(self):
        super().__init__()
        self.torch_style = nn.Embedding(512, 64, padding_idx=3)
        self.hf_style = nn.Embedding(300, 48, padding_idx=0)
        self.hf_style.weight.data.normal_(mean=0.0, std=0.02)
        self.hf_style.weight.data[self.hf_style.padding_idx].zero_()
        self.last_row = nn.Embedding(257, 40, padding_idx=256)
        self.halves = nn.Parameter(torch.empty(64, 32))
        with torch.no_grad():
            self.halves[:32].normal_(0.0, 0.1)
            self.halves[32:].fill_(0.5)
            self.halves[:16].mul_(2.0)


def padded_embeddings():
    return PaddedEmbeddings()


# ---- real constructors at toy sizes (HF `_init_weights`, torch.nn defaults) --------------------------
def families():
    import transformers as TF
    from torch import nn

    small = dict(vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=4,
                 num_key_value_heads=2)
    return {
        "llama": lambda: TF.LlamaForCausalLM(TF.LlamaConfig(**{**small, "num_hidden_layers": 2})),
        "mistral": lambda: TF.MistralForCausalLM(TF.MistralConfig(**small)),
        "qwen2": lambda: TF.Qwen2ForCausalLM(TF.Qwen2Config(**small)),
        "mixtral": lambda: TF.MixtralForCausalLM(TF.MixtralConfig(**small, num_local_experts=2)),
        "gemma2": lambda: TF.Gemma2ForCausalLM(TF.Gemma2Config(**small, head_dim=16)),
        "phi3": lambda: TF.Phi3ForCausalLM(TF.Phi3Config(**small, pad_token_id=0)),
        "gpt2": lambda: TF.GPT2LMHeadModel(TF.GPT2Config(vocab_size=512, n_embd=64, n_layer=2, n_head=4, n_positions=64)),
        "opt": lambda: TF.OPTForCausalLM(TF.OPTConfig(vocab_size=512, hidden_size=64, ffn_dim=128, num_hidden_layers=1,
                                                      num_attention_heads=4, word_embed_proj_dim=64, max_position_embeddings=64)),
        "bert": lambda: TF.BertModel(TF.BertConfig(vocab_size=512, hidden_size=64, num_hidden_layers=1, num_attention_heads=4,
                                                   intermediate_size=128)),
        "t5": lambda: TF.T5Model(TF.T5Config(d_model=64, d_ff=128, num_layers=1, num_heads=4, vocab_size=512)),
        "vit": lambda: TF.ViTModel(TF.ViTConfig(hidden_size=64, num_hidden_layers=1, num_attention_heads=4, intermediate_size=128,
                                                image_size=32, patch_size=8)),
        "convnext": lambda: TF.ConvNextModel(TF.ConvNextConfig(num_channels=3, hidden_sizes=[16, 32], depths=[1, 1], num_stages=2)),
        "lstm": lambda: nn.LSTM(32, 64, num_layers=2),
        "conv_bn": lambda: nn.Sequential(nn.Conv2d(3, 16, 3), nn.BatchNorm2d(16), nn.Linear(10, 10)),
        "mha": lambda: nn.MultiheadAttention(32, 4),
        "transformer": lambda: nn.Transformer(d_model=32, nhead=4, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=64),
    }


FAMILIES = ["llama", "mistral", "qwen2", "mixtral", "gemma2", "phi3", "gpt2", "opt", "bert", "t5", "vit", "lstm", "conv_bn", "mha",
            "transformer"]
# constructors the reference cannot record at all (`tensor.tolist()` on a tensor without storage): engine-only
FAMILIES_ENGINE_ONLY = ["convnext"]


CASES = {
    "linear128": linear128,
    "llama8b_layer": llama8b_layer,
    "llama8b_layer_cast": llama8b_layer_cast,
    "big_inits": big_inits,
    "padded_embeddings": padded_embeddings,
    "init_zoo": init_zoo,
    "tiny_llama": tiny_llama,
    "tiny_gpt2": tiny_gpt2,
    "mlp_stack": mlp_stack,
    "torch_transformer": torch_transformer,
    "clones": clones,
    "cast_variant": cast_variant,
}

DTYPES = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def build(name: str, dtype: str = "fp32", device: str = "cpu"):
    """Builds the case with `dtype` as default dtype and `device` as default device
    (`family_<name>`: one of `families()`)."""
    prev = torch.get_default_dtype()
    torch.set_default_dtype(DTYPES[dtype])
    try:
        with torch.device(device):
            if name.startswith("family_"):
                return families()[name[len("family_"):]]()
            return CASES[name]()
    finally:
        torch.set_default_dtype(prev)
