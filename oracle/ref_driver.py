"""Runs the REAL reference (oracle/_ref) on a named case and saves the materialised tensors.

    python oracle/ref_driver.py --case init_zoo --seed 0 --dtype fp32 --device cpu --out x.pt

TEST INFRASTRUCTURE: a separate process because the reference and torchdistx_b200 both register
fallbacks on DispatchKey::Fake / DispatchKey::DeferredInit and cannot share one.
Also the CPU baseline of bench.py (`--time`: prints the wall time of materialize_module).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases  # noqa: E402
from oracle import ref_torchdistx as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="")
    ap.add_argument("--cases", default="", help="comma list of case:dtype, saved as OUTDIR/case_dtype.pt")
    ap.add_argument("--outdir", default="")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--out", default="")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--stats", action="store_true",
                    help="large tensors (> 2^20 elements) are saved as float64 moments + min/max + an evenly "
                         "strided 2^20-element subsample instead of whole (BASELINE-sized cases)")
    ap.add_argument("--round-to", default="", help="round every floating matrix (dim >= 2) to this dtype first (bf16/fp16)")
    a = ap.parse_args()
    global STATS, ROUND_TO
    STATS, ROUND_TO = a.stats, a.round_to
    if a.threads:
        torch.set_num_threads(a.threads)
    if a.cases:
        for item in a.cases.split(","):
            case, dtype = item.split(":")
            run(case, dtype, a.device, a.seed, os.path.join(a.outdir, f"{case}_{dtype}.pt"), False)
        return
    run(a.case, a.dtype, a.device, a.seed, a.out, a.time)


STATS, ROUND_TO = False, ""
SUBSAMPLE = 1 << 20


def summarize(t: torch.Tensor):
    """What the full-size parity tests need of a tensor too large to ship whole."""
    if ROUND_TO and t.is_floating_point() and t.dim() >= 2:  # (matrices: the RNG-initialised tensors)
        t = t.to(cases.DTYPES[ROUND_TO])
    n = t.numel()
    if not STATS or n <= SUBSAMPLE or not t.is_floating_point():
        return t
    flat = t.reshape(-1)
    d = flat.double()
    stride = n // SUBSAMPLE
    return {"numel": n, "dtype": str(t.dtype), "shape": tuple(t.shape), "mean": d.mean().item(),
            "std": d.std().item(), "min": d.min().item(), "max": d.max().item(),
            "finite": bool(torch.isfinite(d).all()), "stride": stride,
            "sample": flat[: stride * SUBSAMPLE : stride].clone()}


def run(case, dtype, device, seed, out, timed):
    class A:
        pass

    a = A()
    a.case, a.dtype, a.device, a.seed, a.out, a.time = case, dtype, device, seed, out, timed
    # the default dtype must be the same at record and at replay time: the reference resolves
    # `dtype=None` when it replays
    prev = torch.get_default_dtype()
    torch.set_default_dtype(cases.DTYPES[a.dtype])
    m = R.deferred_init(lambda: cases.build(a.case, a.dtype, a.device))
    assert R.is_deferred(m)
    torch.manual_seed(a.seed)
    t0 = time.perf_counter()
    R.materialize_module(m)
    dt = time.perf_counter() - t0
    torch.set_default_dtype(prev)
    assert not R.is_deferred(m)
    if a.out:
        sd = {k: summarize(v.detach().cpu()) for k, v in list(m.named_parameters()) + list(m.named_buffers())}
        torch.save(sd, a.out)
    if a.time:
        n = sum(p.numel() for p in m.parameters())
        print(json.dumps({"case": a.case, "seconds": dt, "params": n, "threads": torch.get_num_threads()}))


if __name__ == "__main__":
    main()
