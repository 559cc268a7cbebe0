"""Host-side cost of `materialize_module(..., shard=(0, W))` on ONE GPU, W = 1, 2, 4, 8.

One process plays rank 0 of W: it plans, allocates and wraps every tensor of the model (that work does
not shrink with W) but writes only 1/W of the bytes, so this shows where the end-to-end time of a
sharded materialise becomes host-bound -- without 8 ranks competing for the box's cores.

    python benchmarks/e2e_shard_probe.py [--model llama3-8b] [--steps 5]
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torchdistx_b200.deferred_init import deferred_init, last_materialize_stats, materialize_module  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--worlds", default="1,2,4,8")
    ap.add_argument("--gc", type=int, default=1, help="0: collect before and disable the Python GC inside every timed call (like timeit)")
    ap.add_argument("--prewarm-ms", type=float, default=0.0,
                    help="keep the GPU busy for this long right before every timed call (diagnostic: how much of the "
                         "end-to-end time is the GPU leaving its idle clocks)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    spin = torch.empty(1 << 24, device=dev)
    for world in [int(w) for w in a.worlds.split(",")]:
        fakes = [deferred_init(bench.build_model, a.model) for _ in range(a.steps + 2)]
        shard = None if world == 1 else (0, world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms, host, acc = [], [], {}
        for i, m in enumerate(fakes):
            torch.cuda.synchronize()
            if not a.gc:
                gc.collect()
                gc.disable()
            if a.prewarm_ms > 0:
                t_end = time.perf_counter() + a.prewarm_ms / 1e3
                while time.perf_counter() < t_end:
                    spin.normal_()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            materialize_module(m, device=dev, shard=shard)
            t1 = time.perf_counter()
            e1.record()
            e1.synchronize()
            if not a.gc:
                gc.enable()
            if i >= 2:  # two warm-up steps (allocator, lazy module loading)
                ms.append(e0.elapsed_time(e1))
                host.append((t1 - t0) * 1e3)
                for k, v in last_materialize_stats().items():
                    acc[k] = acc.get(k, 0) + v
            fakes[i] = None
            del m
        n = len(ms)
        keys = ("traverse_us", "plan_us", "eval_us", "alloc_us", "launch_us", "wrap_us", "first_submit_us",
                "last_submit_us", "submissions", "kernel_launches", "bytes_written")
        print(json.dumps({"model": a.model, "shard": [0, world], "python_gc_inside_timed_call": bool(a.gc), "e2e_ms": round(sum(ms) / n, 3),
                          "api_return_ms": round(sum(host) / n, 3),
                          "host": {k: round(acc.get(k, 0) / n, 1) for k in keys}}), flush=True)


if __name__ == "__main__":
    main()
