"""Writes tests/golden/stream_vectors.json: known-answer vectors of the normative random stream and
of the deterministic sources (include/tdx_init.h), produced by the CPU
restatement oracle/tdx_oracle.c at the commit that introduced them.

    python tests/golden/make_stream_vectors.py

The reference has no vectors of its own for this path (every value comes from the
PyTorch build it runs on), so the restatement is pinned from outside by Random123's and ATen's Philox
vectors (philox_kat.json) and by the distribution / bound tests; THESE vectors pin it against
drifting afterwards: a change to the oracle (and hence to what the kernels are held to, bit for bit,
by tests/test_kernels_gpu.py) must regenerate this file on purpose.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import tdx_oracle as O  # noqa: E402
from torchdistx_b200 import _cabi as C  # noqa: E402

N = 40  # elements per vector: a few Philox blocks, starting inside one (elem_begin is not a multiple of 4 or 8)

TRUNC = [(C.TDX_EPI_ERFINV, 0.0, 0.0), (C.TDX_EPI_MUL, 0.02 * 2 ** 0.5, 0.0), (C.TDX_EPI_ADD, 0.1, 0.0),
         (C.TDX_EPI_CLAMP, -0.04, 0.06)]
INV_FREQ = [(C.TDX_EPI_MUL, 1.0 / 64.0, 0.0), (C.TDX_EPI_RPOW, 500000.0, 0.0), (C.TDX_EPI_RECIP, 0.0, 0.0)]

CASES = [
    # name, kwargs of _cabi.make_desc (dst is irrelevant to the values)
    ("uniform_f32", dict(dtype=C.TDX_F32, src=C.TDX_SRC_UNIFORM, p0=-0.05, p1=0.03)),
    ("uniform_bf16", dict(dtype=C.TDX_BF16, src=C.TDX_SRC_UNIFORM, p0=-0.05, p1=0.03)),
    ("uniform_f16", dict(dtype=C.TDX_F16, src=C.TDX_SRC_UNIFORM, p0=0.0, p1=1.0)),
    ("kaiming_uniform_bf16", dict(dtype=C.TDX_BF16, src=C.TDX_SRC_UNIFORM, p0=-(1 / 4096) ** 0.5, p1=(1 / 4096) ** 0.5)),
    ("normal_f32", dict(dtype=C.TDX_F32, src=C.TDX_SRC_NORMAL, p0=0.0, p1=0.02)),
    ("normal_bf16", dict(dtype=C.TDX_BF16, src=C.TDX_SRC_NORMAL, p0=0.0, p1=0.02)),
    ("normal_f16", dict(dtype=C.TDX_F16, src=C.TDX_SRC_NORMAL, p0=1.0, p1=0.5)),
    ("normal_bf16_from_the_fp32_stream", dict(dtype=C.TDX_BF16, src=C.TDX_SRC_NORMAL, p0=0.0, p1=0.02, algo=C.TDX_ALGO_WIDE32)),
    ("uniform_f16_from_the_fp32_stream", dict(dtype=C.TDX_F16, src=C.TDX_SRC_UNIFORM, p0=-0.3, p1=0.1, algo=C.TDX_ALGO_WIDE32)),
    ("trunc_normal_f32", dict(dtype=C.TDX_F32, src=C.TDX_SRC_UNIFORM, p0=-0.99, p1=0.95, epi=TRUNC)),
    ("trunc_normal_bf16", dict(dtype=C.TDX_BF16, src=C.TDX_SRC_UNIFORM, p0=-0.99, p1=0.95, epi=TRUNC)),
    ("randn_scaled_f32", dict(dtype=C.TDX_F32, src=C.TDX_SRC_NORMAL, p0=0.0, p1=1.0,
                              epi=[(C.TDX_EPI_MUL, 0.02, 0.0), (C.TDX_EPI_ADD, 1.0, 0.0)])),
    ("randn_scaled_bf16", dict(dtype=C.TDX_BF16, src=C.TDX_SRC_NORMAL, p0=0.0, p1=1.0,
                               epi=[(C.TDX_EPI_MUL, 0.02, 0.0), (C.TDX_EPI_ADD, 1.0, 0.0)])),
    ("arange_i64", dict(dtype=C.TDX_I64, src=C.TDX_SRC_IOTA, p0=5, p1=3)),
    ("arange_f32", dict(dtype=C.TDX_F32, src=C.TDX_SRC_IOTA, p0=0, p1=1)),
    ("rotary_inv_freq_f32", dict(dtype=C.TDX_F32, src=C.TDX_SRC_IOTA, p0=0, p1=2, epi=INV_FREQ)),
    ("rotary_inv_freq_bf16", dict(dtype=C.TDX_BF16, src=C.TDX_SRC_IOTA, p0=0, p1=2, epi=INV_FREQ)),
    ("fill_bf16_one", dict(dtype=C.TDX_RAW16, src=C.TDX_SRC_CONST, fill_bits=0x3F80, fill_itemsize=2)),
    ("fill_f32_half", dict(dtype=C.TDX_RAW32, src=C.TDX_SRC_CONST, fill_bits=0x3F000000, fill_itemsize=4)),
    ("fill_i64", dict(dtype=C.TDX_RAW64, src=C.TDX_SRC_CONST, fill_bits=0x0123456789ABCDEF, fill_itemsize=8)),
    ("fill_bool", dict(dtype=C.TDX_RAW8, src=C.TDX_SRC_CONST, fill_bits=1, fill_itemsize=1)),
]
STREAMS = [(0, 0, 0), (0x9E3779B97F4A7C15, 4096, 13), (2 ** 63 + 12345, 2 ** 33, 2 ** 32 + 7)]  # (seed, offset, elem_begin)


def vectors():
    out = []
    for name, kw in CASES:
        rng = kw["src"] in (C.TDX_SRC_UNIFORM, C.TDX_SRC_NORMAL)
        for seed, offset, begin in (STREAMS if rng else STREAMS[:2]):
            if kw["src"] == C.TDX_SRC_IOTA:
                begin = min(begin, 13)  # (index programs must stay exactly representable)
            d = C.make_desc(0, elem_count=N, elem_begin=begin, seed=seed if rng else 0, offset=offset if rng else 0, **kw)
            out.append({"name": name, "seed": seed if rng else 0, "offset": offset if rng else 0, "elem_begin": begin,
                        "elem_count": N, "desc": {k: (list(map(list, v)) if k == "epi" else v) for k, v in kw.items()},
                        "out_hex": O.generate(d).tobytes().hex()})
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stream_vectors.json")
    doc = {"what": __doc__.split("\n\n")[0], "abi_version": C.load().tdx_abi_version(), "vectors": vectors()}
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print(path, len(doc["vectors"]), "vectors")
