"""BASELINE config #5: init-kernel sweep (normal_/uniform_/fill, bf16 + fp32, 1 MB .. 16 GB),
driven through the C ABI (tdx_plan_launch), timed with CUDA events on the launching stream.

    python benchmarks/kernel_sweep.py [--max-gb 4] [--variants all] [--out gpurun_out/sweep.jsonl]

Also times the stock ATen kernels (what the reference's replay dispatches to on a CUDA device,
reference deferred_init.cc:218-220) on the same buffers: "baseline B" of BASELINE.md.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdistx_b200 import _cabi as C  # noqa: E402

PEAK = 6565.8  # MEASURED_PEAKS.json hbm_gbs (copy, read+write)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

VARIANTS = {
    # name: (dtype, torch dtype, src, algo, p0, p1)
    "fill_bf16": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_CONST, 0, 0, 0),
    "fill_f32": (C.TDX_F32, torch.float32, C.TDX_SRC_CONST, 0, 0, 0),
    "uniform_bf16": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_UNIFORM, 0, -0.05, 0.05),
    "uniform_bf16_nolut": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_UNIFORM, C.TDX_ALGO_NOLUT, -0.05, 0.05),
    "uniform_f16": (C.TDX_F16, torch.float16, C.TDX_SRC_UNIFORM, 0, -0.05, 0.05),
    "uniform_bf16_r7": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_UNIFORM, C.TDX_ALGO_R7, -0.05, 0.05),
    "uniform_f32": (C.TDX_F32, torch.float32, C.TDX_SRC_UNIFORM, 0, -0.05, 0.05),
    "normal_bf16_icdf16": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_NORMAL, C.TDX_ALGO_ICDF16, 0.0, 0.02),
    "normal_bf16_icdf16_nolut": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_NORMAL, C.TDX_ALGO_ICDF16 | C.TDX_ALGO_NOLUT, 0.0, 0.02),
    "normal_bf16_icdf16_r7": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_NORMAL, C.TDX_ALGO_ICDF16 | C.TDX_ALGO_R7, 0.0, 0.02),
    "normal_bf16_bm16": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_NORMAL, C.TDX_ALGO_BM16, 0.0, 0.02),
    "normal_bf16_bm16_r7": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_NORMAL, C.TDX_ALGO_BM16 | C.TDX_ALGO_R7, 0.0, 0.02),
    "normal_bf16_bm32": (C.TDX_BF16, torch.bfloat16, C.TDX_SRC_NORMAL, C.TDX_ALGO_BM32, 0.0, 0.02),
    "normal_f32_bm32": (C.TDX_F32, torch.float32, C.TDX_SRC_NORMAL, C.TDX_ALGO_BM32, 0.0, 0.02),
    "normal_f32_bm32_r7": (C.TDX_F32, torch.float32, C.TDX_SRC_NORMAL, C.TDX_ALGO_BM32 | C.TDX_ALGO_R7, 0.0, 0.02),
    "normal_f16_icdf16": (C.TDX_F16, torch.float16, C.TDX_SRC_NORMAL, 0, 0.0, 0.02),
}


def time_cuda(fn, iters, flush):
    """median ms over `iters`, L2 flushed (a > L2 write) before every timed launch"""
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-gb", type=float, default=4.0)
    ap.add_argument("--min-mb", type=float, default=1.0)
    ap.add_argument("--variants", default="all")
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--aten", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    a = ap.parse_args()
    lib = C.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    names = list(VARIANTS) if a.variants == "all" else a.variants.split(",")
    ws_bytes = lib.tdx_init_workspace_bytes(1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    out = open(a.out, "a")
    sizes = []
    b = int(a.min_mb * (1 << 20))
    while b <= int(a.max_gb * (1 << 30)):
        sizes.append(b)
        b *= 4
    if sizes[-1] != int(a.max_gb * (1 << 30)):
        sizes.append(int(a.max_gb * (1 << 30)))
    for nbytes in sizes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        for name in names:
            dt, tdt, src, algo, p0, p1 = VARIANTS[name]
            isz = 4 if dt == C.TDX_F32 else 2
            n = nbytes // isz
            d = C.make_desc(buf.data_ptr(), dtype=dt, src=src, elem_count=n, seed=1234, offset=8,
                            p0=p0, p1=p1, algo=algo, fill_bits=0x3C00 if isz == 2 else 0x3F800000,
                            fill_itemsize=isz)
            arr = (C.TdxInitDesc * 1)(d)
            plan = C.TdxPlan()
            if lib.tdx_plan_upload(arr, 1, ws.data_ptr(), ws_bytes, stream, ctypes.byref(plan)) != 0:
                continue  # e.g. the experimental R7 / BM16 kernels: only in -DTDX_EXPERIMENTAL_ALGOS builds (TDX_INIT_LIB)

            def run():
                C.check(lib.tdx_plan_launch(ctypes.byref(plan), ws.data_ptr(), stream))

            for _ in range(3):
                run()
            torch.cuda.synchronize()
            fl = flush if nbytes < (256 << 20) else None  # big buffers exceed L2 on their own
            med, best = time_cuda(run, a.iters, fl)
            t = buf.view(tdt)
            stats = {}
            if src != C.TDX_SRC_CONST and nbytes <= (1 << 30):
                f = t[: min(n, 1 << 26)].float()
                stats = {"mean": f.mean().item(), "std": f.std().item(), "min": f.min().item(),
                         "max": f.max().item()}
            rec = {"kernel": name, "bytes": nbytes, "ms": med, "ms_best": best,
                   "gbs": nbytes / med / 1e6, "frac_of_peak": nbytes / med / 1e6 / PEAK, **stats}
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
        if a.aten:
            for tdt, tag in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
                t = buf.view(tdt)
                for opname, fn in (("normal_", lambda: t.normal_(0.0, 0.02)),
                                   ("uniform_", lambda: t.uniform_(-0.05, 0.05)),
                                   ("fill_", lambda: t.fill_(1.0))):
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    fl = flush if nbytes < (256 << 20) else None
                    med, best = time_cuda(fn, a.iters, fl)
                    rec = {"kernel": f"aten_{opname}{tag}", "bytes": nbytes, "ms": med, "ms_best": best,
                           "gbs": nbytes / med / 1e6, "frac_of_peak": nbytes / med / 1e6 / PEAK}
                    print(json.dumps(rec), flush=True)
                    out.write(json.dumps(rec) + "\n")
        del buf
        torch.cuda.empty_cache()
    out.close()


if __name__ == "__main__":
    main()
