"""T1 parity at BASELINE tensor sizes (verbatim tolerances).

The engine's CUDA output is compared with the REAL reference (oracle/_ref = pytorch/torchdistx
compiled from its own sources, CPU device, subprocess) running the same program under the same
seed, on

  * `llama8b_layer`      the 234,893,312-parameter sample of Llama-3-8B that bench.py times the
                         reference on (every linear at its real size: 4096x4096, 1024x4096,
                         14336x4096, 4096x14336, 2048x4096 embeddings), bf16;
  * `llama8b_layer_cast` the same model built in fp32 and converted with `.to(torch.bfloat16)`
                         (SURVEY 8d cfg3 variant: TDX_ALGO_WIDE32 descriptors);
  * `big_inits`          one 2^24-element tensor per RNG idiom of BASELINE config 5
                         (normal_, uniform_, kaiming_uniform_, trunc_normal_, randn*s+m), fp32 and bf16.

Per RNG tensor (N elements, sigma = the reference sample's std), SURVEY 8c T1:

    |mean - mean_ref|   <= 5 sigma / sqrt(N)
    |std/std_ref - 1|   <= 5 / sqrt(2N)            (+ 2^-8 for 16-bit outputs: the rounding term)
    hard range          uniform in [from, to), trunc_normal_ in [a, b]   (through the API)
    two-sample KS on an evenly strided 2^20-element subsample, alpha = 1e-3
    no NaN / Inf

Both sides of every comparison are samples, so these bounds are 3.5 standard deviations of the
DIFFERENCE (mean: sd sigma sqrt(2/N); std ratio: sd 1/sqrt(N)); the seeds are fixed, so the outcome
is deterministic.  16-bit outputs are compared with the reference's fp32 sample of the same program
ROUNDED to the dtype (`--round-to`): the reference's own bf16 CPU kernel draws its normals from
8-bit uniforms ($TORCH/include/ATen/native/cpu/DistributionTemplates.h:207-256 with
TransformationHelper.h:84-99, digits = 8: |z| <= 3.33, std 0.7 % low), which is an artefact of
that kernel, not the distribution `normal_` names.  The cast variant needs no such step: its
reference IS an fp32 sample rounded by `.to(bfloat16)`.

Deterministic tensors stay T0: norm weights bit-exact; the rotary `inv_freq` buffers (float
arithmetic through `pow`, replayed by ATen's CUDA kernels) to 4 ulp of the reference's CPU result.
"""
import math
import os
import subprocess
import sys

import pytest
import torch

from oracle import cases
from torchdistx_b200.deferred_init import deferred_init, last_materialize_stats, materialize_module

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 11
SUBSAMPLE = 1 << 20

# (case, dtype the ENGINE builds in, dtype the REFERENCE runs in, --round-to)
RUNS = {
    "llama_bf16": ("llama8b_layer", "bf16", "fp32", "bf16"),
    "llama_cast": ("llama8b_layer_cast", "fp32", "fp32", ""),
    "big_fp32": ("big_inits", "fp32", "fp32", ""),
    "big_bf16": ("big_inits", "bf16", "fp32", "bf16"),
}


def run_reference(tmp, key):
    case, _, ref_dtype, round_to = RUNS[key]
    path = os.path.join(tmp, f"{key}.pt")
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_driver.py"), "--case", case, "--dtype", ref_dtype,
           "--seed", str(SEED), "--stats", "--out", path]
    if round_to:
        cmd += ["--round-to", round_to]
    subprocess.run(cmd, check=True, cwd=ROOT, timeout=300)
    return torch.load(path)


@pytest.fixture(scope="module")
def reference(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("t1_fullsize"))
    cache = {}

    def get(key):
        if key not in cache:
            cache[key] = run_reference(tmp, key)
        return cache[key]

    return get


def build(key):
    case, dtype, _, _ = RUNS[key]
    m = deferred_init(lambda: cases.build(case, dtype, "cuda"))
    torch.manual_seed(SEED)
    materialize_module(m)
    return m


def check_t1(name, x, ref, sixteen_bit):
    """SURVEY 8c T1 for one RNG tensor: `x` the engine's CUDA tensor, `ref` the reference's summary."""
    from scipy import stats

    n = x.numel()
    xd = x.detach().double().flatten()
    assert bool(torch.isfinite(xd).all()), name
    if isinstance(ref, dict):
        assert ref["finite"] and ref["numel"] == n and tuple(ref["shape"]) == tuple(x.shape), name
        r_mean, r_std, stride = ref["mean"], ref["std"], ref["stride"]
        r_sample = ref["sample"].double().numpy()
        x_sample = xd[: stride * SUBSAMPLE: stride].cpu().numpy()
    else:
        rd = ref.double().flatten()
        r_mean, r_std = rd.mean().item(), rd.std().item()
        r_sample, x_sample = rd.numpy(), xd.cpu().numpy()
    mean, std = xd.mean().item(), xd.std().item()
    assert abs(mean - r_mean) <= 5 * r_std / math.sqrt(n), (name, mean, r_mean, r_std, n)
    slack = 2.0 ** -8 if sixteen_bit else 0.0
    assert abs(std / r_std - 1) <= 5 / math.sqrt(2 * n) + slack, (name, std, r_std, n)
    if n >= 4096:
        ks = stats.ks_2samp(x_sample, r_sample)
        assert ks.pvalue > 1e-3, (name, ks)
    return mean, std


@pytest.mark.parametrize("key", ["llama_bf16", "llama_cast"])
def test_llama3_8b_sample_against_reference(key, reference):
    ref = reference(key)
    m = build(key)
    st = last_materialize_stats()
    mine = dict(list(m.named_parameters()) + list(m.named_buffers()))
    assert set(mine) <= set(ref)
    n_rng = n_big = 0
    total = 0
    for k, t in mine.items():
        r = ref[k]
        total += t.numel() * t.element_size()
        if t.dim() < 2:  # deterministic programs: norm weights (ones) and the rotary inv_freq buffers
            # (the bf16 model's reference runs in fp32: a constant program commutes with the cast)
            assert not isinstance(r, dict), k
            x, rr = t.detach().cpu(), r.to(t.dtype)
            if "inv_freq" in k:
                # arange -> div -> pow -> reciprocal, replayed by ATen on the GPU: CUDA's powf and the
                # CPU's vectorised pow are both faithful to ~1 ulp but not to each other's last bit
                # (T0 is stated for zeros/ones/index ops; this is float arithmetic: 4 ulp)
                torch.testing.assert_close(x, rr, rtol=4 * 2.0 ** -23, atol=0.0)
            else:
                assert torch.equal(x, rr), k
            continue
        r_dtype = r["dtype"] if isinstance(r, dict) else str(r.dtype)
        assert t.is_cuda and str(t.dtype) == r_dtype, (k, t.dtype, r_dtype)
        n_rng += 1
        n_big += t.numel() >= (1 << 22)
        _, std = check_t1(k, t, r, sixteen_bit=True)
        assert abs(std - 0.02) < 2e-4, (k, std)  # HF Llama: normal_(0, initializer_range = 0.02)
        assert t.detach().abs().max().item() <= 0.02 * 8.5, k  # the generator reaches +-7.9 sigma, no further
    assert n_rng >= 9 and n_big >= 7, (n_rng, n_big)
    assert sum(p.numel() for p in m.parameters()) == 234_893_312
    # everything large went through the fused kernels (inv_freq is the only generic program)
    assert st["bytes_written"] >= 0.999 * total, (st, total)


@pytest.mark.parametrize("key", ["big_fp32", "big_bf16"])
def test_single_2p24_tensors_against_reference(key, reference):
    ref = reference(key)
    m = build(key)
    sixteen = key.endswith("bf16")
    st = last_materialize_stats()
    assert st["fused_tensors"] == 5 and st["generic_ops"] == 0, st
    got = {}
    for k, t in m.named_parameters():
        assert t.numel() == 1 << 24 and isinstance(ref[k], dict), k
        got[k] = check_t1(k, t, ref[k], sixteen)
    # hard range checks through the API: uniform in [from, to), trunc in [a, b]
    dt = torch.bfloat16 if sixteen else torch.float32
    lim = lambda v: torch.tensor(v, dtype=dt).item()  # uniform_ rounds its bounds to the tensor dtype
    u = m.uniform.detach()
    assert u.min().item() >= lim(-0.05) and u.max().item() < lim(0.03), (u.min().item(), u.max().item())
    k_ = m.kaiming.detach()
    bound = math.sqrt(3.0) * math.sqrt(2.0 / (1 + 5.0)) / math.sqrt(cases.BIG)  # calculate_gain('leaky_relu', sqrt 5)
    assert k_.min().item() >= lim(-bound) and k_.max().item() < lim(bound), (k_.min().item(), k_.max().item(), bound)
    tr = m.trunc.detach()
    assert tr.min().item() >= lim(-0.04) and tr.max().item() <= lim(0.04), (tr.min().item(), tr.max().item())
    # and the moments the programs name
    assert abs(got["normal"][1] - 0.02) < 1e-4 and abs(got["scaled"][0] - 1.0) < 1e-3
    assert abs(got["uniform"][0] + 0.01) < 1e-4 and abs(got["uniform"][1] - 0.08 / math.sqrt(12)) < 1e-4
