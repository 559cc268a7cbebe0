"""Per-CTA timeline of the table kernel on the bench workload (measurement build only).

    benchmarks/build_variants.sh "timeline:-DTDX_LUT_TIMELINE"
    python benchmarks/lut_timeline.py [--as-rank-of 8] [--model llama3-8b]

Materialises the model once through the public API (product library) to obtain the descriptor
table, then launches the same plan through the instrumented build of libtdx_init.so and prints where
the CTAs of one launch spend their time (globaltimer, ns): table build, grab boundaries, tail.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--as-rank-of", type=int, default=1)
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--lib", default=os.path.join(ROOT, "benchmarks", "_variants", "libtdx_timeline.so"))
    a = ap.parse_args()
    import torch

    from torchdistx_b200 import _cabi as C

    C.LIB_PATH = a.lib  # ctypes calls below go to the instrumented build; the engine keeps its own
    import bench
    from torchdistx_b200.deferred_init import deferred_init, last_descriptors, materialize_module

    dev = torch.device("cuda:0")
    m = deferred_init(bench.build_model, a.model)
    materialize_module(m, device=dev, shard=(0, a.as_rank_of) if a.as_rank_of > 1 else None)
    descs = last_descriptors()
    lib = C.load()
    lib.tdx_debug_lut_timeline.argtypes = [ctypes.c_void_p]
    lib.tdx_debug_lut_timeline.restype = ctypes.c_int
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    buf = torch.zeros(sms * 16, dtype=torch.int64, device=dev)
    assert lib.tdx_debug_lut_timeline(buf.data_ptr()) == 0
    ws = torch.empty(lib.tdx_init_workspace_bytes(len(descs)), dtype=torch.uint8, device=dev)
    plan = C.TdxPlan()
    stream = torch.cuda.current_stream().cuda_stream
    C.check(lib.tdx_plan_upload(descs, len(descs), ws.data_ptr(), ws.numel(), stream, ctypes.byref(plan)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        C.check(lib.tdx_plan_launch(ctypes.byref(plan), ws.data_ptr(), stream))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        C.check(lib.tdx_plan_launch(ctypes.byref(plan), ws.data_ptr(), stream))
    e1.record()
    torch.cuda.synchronize()
    t = buf.cpu().view(sms, 16).double()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    us = lambda x: float(x) / 1e3

    def stat(col):
        return {"min": round(us(col.min()), 2), "mean": round(us(col.mean()), 2), "max": round(us(col.max()), 2)}

    out = {
        "model": a.model, "as_rank_of": a.as_rank_of, "ctas": int(t.shape[0]),
        "launch_ms_events_incl_fills": round(e0.elapsed_time(e1) / 10, 4),
        "bytes": int(sum(bench.desc_bytes(C, d) for d in descs)),
        "enter_us": stat(t[:, 0] - t0), "first_grab_known_us": stat(t[:, 1] - t[:, 0]),
        "first_table_done_us": stat(t[:, 2] - t[:, 0]), "last_grab_done_us": stat(t[:, 3] - t0),
        "exit_us": stat(t[:, 7] - t0), "grabs": stat(t[:, 4] * 1e3), "tiles": stat(t[:, 8] * 1e3),
        "barrier_wait_thread0_us": stat(t[:, 5]), "grab_setup_us": stat(t[:, 6]),  # (barriers: end of the share + one per tail grab)
        "table_builds": stat(t[:, 9] * 1e3), "table_build_us": stat(t[:, 10]),
        "share_done_us": stat(t[t[:, 13] > 0][:, 13] - t0) if bool((t[:, 13] > 0).any()) else None,
        "busy_us": stat(t[:, 3] - t[:, 0]),
        # when a CTA asked the work counter for the first time (one grab before its pre-assigned share ends)
        "first_dynamic_request_us": stat(t[t[:, 12] > 0][:, 12] - t0) if bool((t[:, 12] > 0).any()) else None,
        "env": {k: v for k, v in os.environ.items() if k.startswith("TDX_LUT_")},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
