"""Runs the REAL reference (oracle/_ref) over the random scripts of oracle/fuzz_programs.py and saves
what it materialises (TEST INFRASTRUCTURE; a separate process for the same reason as ref_driver.py).

    python oracle/ref_fuzz_driver.py --lo 0 --hi 300 --out ref.pt

Per seed: `deferred_init(LinkedHolder, script)`, `torch.manual_seed(seed)`, `materialize_module` --
CPU device, so the reference's op-by-op replay through ATen (deferred_init.cc:506-667) defines every
bit, random draws included.  A script the reference refuses is saved as its error text."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fuzz_programs as P  # noqa: E402
from oracle import ref_torchdistx as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lo", type=int, default=0)
    ap.add_argument("--hi", type=int, default=100)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    out = {}
    for seed in range(a.lo, a.hi):
        progs, links = P.differential_script(seed)
        try:
            m = R.deferred_init(P.LinkedHolder, progs, links)
            torch.manual_seed(seed)
            R.materialize_module(m)
            out[seed] = {k: v.detach().clone() for k, v in m.named_parameters()}
        except Exception as e:  # noqa: BLE001  (whatever the reference says about this script)
            out[seed] = f"{type(e).__name__}: {e}"
    torch.save(out, a.out)


if __name__ == "__main__":
    main()
