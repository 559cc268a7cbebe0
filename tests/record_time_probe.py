"""(Not a test: a probe that lives under tests/ because it runs the compiled reference, which only tests/,
smoke() and bench.py may do.)

Recorder overhead: seconds per `deferred_init(model)` for this engine and for the
compiled reference (oracle/_ref), each in a process of its own, on the CPU -- no GPU involved.

    python tests/record_time_probe.py [--models llama3-8b,llama3-70b,gpt2-xl] [--reps 5]

Prints one JSON line per (engine, model): the times of `reps` recordings after one warm-up, the
recorded op count where the engine exposes it, and -- as the floor -- the same constructor on the
meta device with no recording at all."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r"""
import json, sys, time, torch
sys.path.insert(0, {root!r})
which, model, reps = {which!r}, {model!r}, {reps}
if which == "reference":
    from oracle import ref_torchdistx as R
    deferred_init = R.deferred_init
elif which == "engine":
    from torchdistx_b200.deferred_init import deferred_init
import bench
torch.set_default_dtype({{"bf16": torch.bfloat16, "fp32": torch.float32, "fp32->bf16": torch.float32}}[bench.MODELS[model][2]])
times = []
for i in range(reps + 1):
    t0 = time.perf_counter()
    if which == "meta":
        with torch.device("meta"):
            m = bench.build_model(model)
    else:
        m = deferred_init(lambda: bench.build_model(model))
    times.append(time.perf_counter() - t0)
    n = sum(p.numel() for p in m.parameters())
    del m
print(json.dumps({{"engine": which, "model": model, "params": n, "seconds": [round(t, 4) for t in times[1:]]}}))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="llama3-8b,llama3-70b,gpt2-xl")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    for model in a.models.split(","):
        for which in ("meta", "engine", "reference"):
            code = SNIPPET.format(root=ROOT, which=which, model=model, reps=a.reps)
            out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=1800)
            line = out.stdout.strip().splitlines()[-1] if out.returncode == 0 and out.stdout.strip() else json.dumps(
                {"engine": which, "model": model, "error": out.stderr.strip()[-300:]})
            print(line, flush=True)


if __name__ == "__main__":
    main()
