"""T0' parity: for CPU-device tensors the engine replays through ATen and
must equal the REAL reference (oracle/_ref, run in a subprocess) bit for bit, RNG included --
same mt19937 stream, same materialize order, dead ops included."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import cases
from torchdistx_b200.deferred_init import deferred_init, materialize_module

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_tdx_ref.so")

pytestmark = pytest.mark.skipif(
    not os.path.exists(REF_SO) and not os.path.isdir("/root/reference"),
    reason="reference oracle not built (python oracle/build_ref.py)")


CASES = [
    ("linear128", "fp32"), ("init_zoo", "fp32"), ("init_zoo", "bf16"), ("mlp_stack", "fp32"),
    ("tiny_llama", "fp32"), ("tiny_llama", "bf16"), ("tiny_gpt2", "fp32"),
    ("torch_transformer", "fp32"), ("clones", "fp32"), ("cast_variant", "fp32"),
] + [(f"family_{name}", "fp32") for name in cases.FAMILIES] + [("family_llama", "bf16"), ("family_gpt2", "bf16")]
# (families: real HF / torch.nn constructors at toy sizes -- Llama, Mistral, Qwen2, Mixtral, Gemma-2, Phi-3,
#  GPT-2, OPT, BERT, T5, ViT, LSTM, Conv+BatchNorm, MultiheadAttention, nn.Transformer)
SEED = 5


@pytest.fixture(scope="module")
def reference_outputs(tmp_path_factory):
    """One oracle subprocess materialises every case with the real reference."""
    outdir = tmp_path_factory.mktemp("ref")
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), "--cases",
                    ",".join(f"{c}:{d}" for c, d in CASES), "--seed", str(SEED), "--outdir", str(outdir)],
                   check=True, cwd=ROOT, timeout=300)
    return outdir


@pytest.mark.parametrize("case,dtype", CASES)
def test_cpu_materialize_equals_reference_bit_exact(case, dtype, reference_outputs):
    ref = torch.load(reference_outputs / f"{case}_{dtype}.pt")
    torch.set_default_dtype(cases.DTYPES[dtype])
    try:
        m = deferred_init(lambda: cases.build(case, dtype))
        torch.manual_seed(SEED)
        materialize_module(m)
    finally:
        torch.set_default_dtype(torch.float32)
    mine = dict(list(m.named_parameters()) + list(m.named_buffers()))
    # tied parameters stay ONE object here (named_parameters de-duplicates them); the oracle
    # driver re-wraps per slot, so it may list more names
    assert set(mine) <= set(ref) and len(mine) >= len(ref) - 2
    for k, t in mine.items():
        r = ref[k]
        assert t.dtype == r.dtype and t.shape == r.shape, k
        assert torch.equal(t.detach(), r), k
