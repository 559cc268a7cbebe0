"""GPU parity, kernel level: libtdx_init driven through its C ABI (raw device pointers) against
the CPU oracle (oracle/tdx_oracle.c) on the same descriptors.

Bars (stated here, enforced below):
  * constant fills, uniform (all dtypes), every rounding/epilogue/sharding/ragged-edge case: BIT-EXACT;
  * normal fp32 (Box-Muller through MUFU lg2/sqrt/sin/cos vs libm), with z = (x - mean)/std:
      |gpu - oracle| <= std * min(1e-3, 4e-7/|z| + 2e-6 (1 + |z|)) + 2 ulp(x)
    (lg2.approx has 2^-22 ABSOLUTE error, so the radius sqrt(-2 ln u) of a draw with u ~ 1 carries
    an error ~2e-7/r: large relative to a tiny radius, irrelevant to the distribution);
  * normal bf16/fp16 (inverse CDF through MUFU lg2 vs libm log2f): identical bits except where the
    fp32 value sits within MUFU error of a rounding boundary: <= 0.2 % of elements may differ, and
    then by exactly 1 ulp of the output dtype.
"""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import tdx_oracle as O
from torchdistx_b200 import _cabi as C

pytestmark = pytest.mark.gpu

TORCH_DT = {C.TDX_F32: torch.float32, C.TDX_BF16: torch.bfloat16, C.TDX_F16: torch.float16}
NP_BITS = {C.TDX_F32: np.uint32, C.TDX_BF16: np.uint16, C.TDX_F16: np.uint16}


def run_descs(descs, bufs):
    lib = C.load()
    n = len(descs)
    ws_bytes = lib.tdx_init_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    launches = C.launch(descs, ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return launches


def gpu_bits(t, dtype):
    return t.view(torch.int32 if dtype == C.TDX_F32 else torch.int16).cpu().numpy().view(NP_BITS[dtype])


def as_float(bits, dtype):
    if dtype == C.TDX_F32:
        return bits.view(np.float32).astype(np.float64)
    if dtype == C.TDX_BF16:
        return (bits.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    return bits.view(np.float16).astype(np.float64)


SIZES = [1, 7, 8, 9, 1023, 1024, 1025, 8191, 100003, (1 << 20) + 5]


@pytest.mark.parametrize("dtype", [C.TDX_F32, C.TDX_BF16, C.TDX_F16])
@pytest.mark.parametrize("n", SIZES)
def test_uniform_bit_exact(dtype, n):
    buf = torch.zeros(n + 16, dtype=TORCH_DT[dtype], device="cuda")
    d = C.make_desc(buf.data_ptr(), dtype=dtype, src=C.TDX_SRC_UNIFORM, elem_count=n, seed=42, offset=12,
                    p0=-0.0625, p1=0.125)
    assert run_descs([d], [buf]) == 1
    got = gpu_bits(buf, dtype)
    assert np.array_equal(got[:n], O.generate(d))
    assert not got[n:].any()  # nothing written past the end


@pytest.mark.parametrize("dtype", [C.TDX_F32, C.TDX_BF16, C.TDX_F16])
@pytest.mark.parametrize("n", [9, 1025, 100003, (1 << 20) + 5])
def test_normal_within_stated_tolerance(dtype, n):
    mean, std = 0.25, 0.02
    buf = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
    d = C.make_desc(buf.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_count=n, seed=7, offset=100,
                    p0=mean, p1=std)
    run_descs([d], [buf])
    got, exp = gpu_bits(buf, dtype), O.generate(d)
    if dtype == C.TDX_F32:
        g, e = as_float(got, dtype), as_float(exp, dtype)
        z = np.abs(e - mean) / std
        tol = std * np.minimum(1e-3, 4e-7 / np.maximum(z, 1e-30) + 2e-6 * (1 + z)) + 2 * np.spacing(np.abs(e).astype(np.float32))
        assert np.all(np.abs(g - e) <= tol), float((np.abs(g - e) / tol).max())
        assert np.mean(np.abs(g - e) <= 2e-6 * std * (1 + z) + 2 * np.spacing(np.abs(e).astype(np.float32))) > 0.99
    else:
        diff = got.astype(np.int64) - exp.astype(np.int64)
        assert np.abs(diff).max() <= 1
        assert (diff != 0).mean() <= 0.002 + 2.0 / n


@pytest.mark.parametrize("dtype", [C.TDX_BF16, C.TDX_F32])
def test_tail_refinement_matches_oracle(dtype):
    # 2^24 elements: ~256 elements of the 16-bit normal take the k == 0 refinement path
    n = 1 << 24
    buf = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
    d = C.make_desc(buf.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_count=n, seed=1, offset=0, p1=1.0)
    run_descs([d], [buf])
    x = buf.float()
    assert torch.isfinite(x).all()
    if dtype == C.TDX_BF16:
        far = (x.abs() > 4.17).nonzero().flatten().cpu().numpy()
        assert 150 <= len(far) <= 400
        # check exactly those elements against the oracle
        got = gpu_bits(buf, dtype)
        for g in far[:64]:
            one = C.make_desc(0, dtype=dtype, src=C.TDX_SRC_NORMAL, elem_begin=int(g), elem_count=1, seed=1,
                              offset=0, p1=1.0)
            assert abs(int(got[g]) - int(O.generate(one)[0])) <= 1
    assert abs(x.double().std().item() - 1) < 1e-3 and abs(x.double().mean().item()) < 2e-3


@pytest.mark.parametrize("itemsize,bits", [(1, 0x01), (2, 0x3F80), (4, 0x3F800000), (8, 0x0123456789ABCDEF)])
@pytest.mark.parametrize("n,shift", [(1, 0), (5, 3), (1000, 1), (65537, 7), ((1 << 20) + 3, 0)])
def test_fill_bit_exact_any_alignment(itemsize, bits, n, shift):
    raw = torch.zeros((n + shift + 16) * itemsize, dtype=torch.uint8, device="cuda")
    dt = {1: C.TDX_RAW8, 2: C.TDX_RAW16, 4: C.TDX_RAW32, 8: C.TDX_RAW64}[itemsize]
    d = C.make_desc(raw.data_ptr() + shift * itemsize, dtype=dt, src=C.TDX_SRC_CONST, elem_count=n,
                    fill_bits=bits, fill_itemsize=itemsize)
    run_descs([d], [raw])
    got = raw.cpu().numpy()
    exp = np.zeros_like(got)
    exp[shift * itemsize:(shift + n) * itemsize] = np.frombuffer(
        int(bits).to_bytes(itemsize, "little") * n, dtype=np.uint8)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("dtype,src", [(C.TDX_BF16, C.TDX_SRC_NORMAL), (C.TDX_F32, C.TDX_SRC_NORMAL),
                                       (C.TDX_BF16, C.TDX_SRC_UNIFORM), (C.TDX_F32, C.TDX_SRC_UNIFORM)])
def test_shards_concatenate_to_the_unsharded_tensor_bit_exact(dtype, src):
    n = 300007
    full = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
    run_descs([C.make_desc(full.data_ptr(), dtype=dtype, src=src, elem_count=n, seed=99, offset=40, p0=0.0, p1=0.5)], [full])
    cuts = [0, 1, 4096, 4099, 123457, 200000, n]  # aligned and unaligned boundaries, ragged sizes
    parts, descs = [], []
    for b, e in zip(cuts[:-1], cuts[1:]):
        t = torch.zeros(e - b, dtype=TORCH_DT[dtype], device="cuda")
        parts.append(t)
        descs.append(C.make_desc(t.data_ptr(), dtype=dtype, src=src, elem_begin=b, elem_count=e - b, seed=99,
                                 offset=40, p0=0.0, p1=0.5))
    assert run_descs(descs, parts) == 1  # one launch for all shards of one kernel family
    assert torch.equal(torch.cat(parts).view(torch.uint8), full.view(torch.uint8))


def test_epilogue_trunc_normal_chain_bit_exact_in_uniform_part():
    # trunc_normal_(mean=.1, std=.02, a=-.04, b=.06) as recorded: uniform_(2l-1, 2u-1) -> erfinv_ ->
    # mul_(std*sqrt2) -> add_(mean) -> clamp_(a, b)
    import math
    mean, std, a, b = 0.1, 0.02, -0.04, 0.06
    lo = 2 * (0.5 * (1 + math.erf((a - mean) / std / math.sqrt(2)))) - 1
    hi = 2 * (0.5 * (1 + math.erf((b - mean) / std / math.sqrt(2)))) - 1
    n = 200001
    for dtype in (C.TDX_F32, C.TDX_BF16):
        buf = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
        epi = [(C.TDX_EPI_ERFINV,), (C.TDX_EPI_MUL, std * math.sqrt(2)), (C.TDX_EPI_ADD, mean),
               (C.TDX_EPI_CLAMP, a, b)]
        d = C.make_desc(buf.data_ptr(), dtype=dtype, src=C.TDX_SRC_UNIFORM, elem_count=n, seed=3, offset=8,
                        p0=lo, p1=hi, epi=epi)
        run_descs([d], [buf])
        g, e = as_float(gpu_bits(buf, dtype), dtype), as_float(O.generate(d), dtype)
        ra, rb = (float(torch.tensor(v, dtype=TORCH_DT[dtype])) for v in (a, b))  # clamp_ sees bounds in dtype
        assert g.min() >= ra - 1e-7 and g.max() <= rb + 1e-7
        # erfinvf (CUDA libm, ~2 ulp) vs the oracle's double-precision inverse: 1 ulp of slack
        tol = 3e-7 if dtype == C.TDX_F32 else 2 ** -8 * 0.1
        assert np.abs(g - e).max() <= tol + 1e-6 * np.abs(e).max()


def test_many_descriptors_one_launch_per_family_and_empty_descriptors():
    descs, bufs = [], []
    for i in range(40):
        n = [0, 1, 33, 4096, 70001][i % 5]
        t = torch.zeros(max(n, 1), dtype=torch.bfloat16, device="cuda")
        bufs.append((t, n))
        descs.append(C.make_desc(t.data_ptr(), dtype=C.TDX_BF16, src=C.TDX_SRC_NORMAL if i % 2 else C.TDX_SRC_UNIFORM,
                                 elem_count=n, seed=5, offset=4 * i, p0=0.0, p1=1.0))
    assert run_descs(descs, bufs) == 2
    for (t, n), d in zip(bufs, descs):
        if n:
            diff = gpu_bits(t, C.TDX_BF16)[:n].astype(np.int64) - O.generate(d).astype(np.int64)
            assert np.abs(diff).max() <= 1


def test_determinism_and_seed_sensitivity():
    n = 1 << 16
    a, b, c = (torch.zeros(n, dtype=torch.bfloat16, device="cuda") for _ in range(3))
    mk = lambda t, seed: C.make_desc(t.data_ptr(), dtype=C.TDX_BF16, src=C.TDX_SRC_NORMAL, elem_count=n, seed=seed, offset=0)
    run_descs([mk(a, 1)], [a]); run_descs([mk(b, 1)], [b]); run_descs([mk(c, 2)], [c])
    assert torch.equal(a, b) and (a != c).float().mean() > 0.9


@pytest.mark.parametrize("dtype", [C.TDX_BF16, C.TDX_F16])
def test_table_driven_uniform_is_bit_identical_to_the_direct_kernel_and_the_oracle(dtype):
    """16-bit uniforms of large descriptors go through the same table kernel (table entry k =
    the direct kernel's value for half-word k); different bounds force table rebuilds."""
    sizes = [(1 << 26) + 9, (1 << 21) + 12345, 1 << 20, 5000]
    bounds = [(-0.05, 0.05), (0.0, 1.0), (-3.0, -1.0), (2.0, 2.5)]
    outs = {}
    for flag in (0, C.TDX_ALGO_NOLUT):
        bufs, descs = [], []
        for i, (n, (lo, hi)) in enumerate(zip(sizes, bounds)):
            t = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
            bufs.append(t)
            descs.append(C.make_desc(t.data_ptr(), dtype=dtype, src=C.TDX_SRC_UNIFORM, elem_count=n, seed=91,
                                     offset=4096 * i, p0=lo, p1=hi, algo=flag))
        launches = run_descs(descs, bufs)
        assert launches == (2 if flag == 0 else 1)  # table kernel + direct kernel (small descriptor)
        outs[flag] = (bufs, descs)
    for a, b in zip(outs[0][0], outs[C.TDX_ALGO_NOLUT][0]):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    for t, d, (lo, hi) in zip(outs[0][0], outs[0][1], bounds):
        head = C.make_desc(0, dtype=dtype, src=C.TDX_SRC_UNIFORM, elem_count=1 << 12, seed=91,
                           offset=d.philox_offset, p0=lo, p1=hi)
        assert np.array_equal(gpu_bits(t[:1 << 12], dtype), O.generate(head))
        f = t.float()
        lo_r, hi_r = (float(torch.tensor(v, dtype=TORCH_DT[dtype])) for v in (lo, hi))  # bounds as the dtype sees them
        assert float(f.min()) >= lo_r and float(f.max()) <= hi_r


@pytest.mark.parametrize("dtype", [C.TDX_BF16, C.TDX_F16])
def test_table_kernel_with_epilogues_is_bit_identical_to_the_direct_kernel(dtype):
    """Descriptors with epilogue steps (trunc_normal_, randn * s + m) are table-driven too: the table
    holds the final value of every half-word.  Same bits as the direct kernel, k == 0 tails included."""
    std, mean, a, b = 0.02, 0.0, -0.04, 0.04
    lo = math.erf((a - mean) / std / math.sqrt(2.0))
    hi = math.erf((b - mean) / std / math.sqrt(2.0))
    trunc = [(C.TDX_EPI_ERFINV,), (C.TDX_EPI_MUL, std * math.sqrt(2)), (C.TDX_EPI_ADD, mean), (C.TDX_EPI_CLAMP, a, b)]
    affine = [(C.TDX_EPI_MUL, 0.02), (C.TDX_EPI_ADD, 1.0)]
    specs = [  # (src, p0, p1, epilogue, elements)
        (C.TDX_SRC_UNIFORM, lo, hi, trunc, (1 << 26) + 9),
        (C.TDX_SRC_NORMAL, 0.0, 1.0, affine, (1 << 26) + 77),
        (C.TDX_SRC_NORMAL, 0.0, 1.0, [(C.TDX_EPI_MUL, 0.5)], (1 << 21) + 3),   # same source, other epilogue: other table
        (C.TDX_SRC_UNIFORM, lo, hi, trunc, 70001),                            # small: direct kernel either way
    ]
    outs = {}
    for flag in (0, C.TDX_ALGO_NOLUT):
        bufs, descs = [], []
        for i, (src, p0, p1, epi, n) in enumerate(specs):
            t = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
            bufs.append(t)
            descs.append(C.make_desc(t.data_ptr(), dtype=dtype, src=src, elem_count=n, seed=17, offset=10000 * i,
                                     p0=p0, p1=p1, epi=epi, algo=flag))
        launches = run_descs(descs, bufs)
        assert launches == (3 if flag == 0 else 2)  # two table kernels + the direct uniform (small descriptor) / two direct ones
        outs[flag] = bufs
    for x, y in zip(outs[0], outs[C.TDX_ALGO_NOLUT]):
        assert torch.equal(x.view(torch.int16), y.view(torch.int16))
    t = outs[0][0].float()
    assert float(t.min()) >= a - 1e-3 and float(t.max()) <= b + 1e-3 and abs(float(t.std()) - 0.0176) < 2e-3  # trunc at +-2 sigma
    g = outs[0][1].float()
    assert abs(float(g.mean()) - 1.0) < 1e-3 and abs(float(g.std()) - 0.02) < 2e-3 and torch.isfinite(g).all()


def test_table_kernel_with_mixed_seeds_and_across_a_2_32_block_boundary():
    """The table kernel takes its Philox round keys from kernel parameters when the launch shares
    one seed and from the descriptors otherwise; inside a tile it treats counter.y (block >> 32) as
    uniform and leaves a tile that straddles a 2^32-block boundary to the generic path.  All of it
    must give the bits of the direct kernel."""
    dtype = C.TDX_BF16
    n = (1 << 26) + 777
    for seeds in ((11, 11, 11), (11, 12, 13)):
        outs = {}
        for flag in (0, C.TDX_ALGO_NOLUT):
            bufs, descs = [], []
            for i, seed in enumerate(seeds):
                t = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
                bufs.append(t)
                descs.append(C.make_desc(t.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_count=n, seed=seed,
                                         offset=64 * i, p0=0.0, p1=0.02, algo=C.TDX_ALGO_ICDF16 | flag))
            run_descs(descs, bufs)
            outs[flag] = bufs
        for a, b in zip(outs[0], outs[C.TDX_ALGO_NOLUT]):
            assert torch.equal(a.view(torch.int16), b.view(torch.int16))
        if len(set(seeds)) > 1:
            assert not torch.equal(outs[0][0], outs[0][1])
    # a shard of a (virtual) 2^35+ element tensor: block indices cross 2^32 in the middle
    m = 1 << 27
    begin = (1 << 35) - (1 << 26) - 24
    outs = []
    for flag in (0, C.TDX_ALGO_NOLUT):
        t = torch.zeros(m, dtype=torch.bfloat16, device="cuda")
        run_descs([C.make_desc(t.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_begin=begin, elem_count=m,
                               seed=3, offset=16, p1=1.0, algo=C.TDX_ALGO_ICDF16 | flag)], [t])
        outs.append(t)
    assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16))
    # spot-check both sides of the boundary against the oracle
    for off in (0, (1 << 26) + 24 - 64):
        d = C.make_desc(0, dtype=dtype, src=C.TDX_SRC_NORMAL, elem_begin=begin + off, elem_count=4096, seed=3,
                        offset=16, p1=1.0)
        diff = gpu_bits(outs[0][off:off + 4096], dtype).astype(np.int64) - O.generate(d).astype(np.int64)
        assert np.abs(diff).max() <= 1


@pytest.mark.parametrize("dtype", [C.TDX_BF16, C.TDX_F16])
def test_table_driven_normal_is_bit_identical_to_the_direct_kernel(dtype):
    """Large descriptors take the shared-memory-table kernel; TDX_ALGO_NOLUT forces the direct one.
    Mixed sizes in one launch, two different (mean, std) so that CTAs rebuild their table."""
    # the table kernel is used when a launch holds >= 2^26 table-eligible elements
    sizes = [(1 << 27) + 9, 1 << 20, (1 << 21) + 12345, 5000]
    outs = {}
    for flag in (0, C.TDX_ALGO_NOLUT):
        bufs, descs = [], []
        for i, n in enumerate(sizes):
            t = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
            bufs.append(t)
            descs.append(C.make_desc(t.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_count=n, seed=77,
                                     offset=1000 * i, p0=0.0 if i % 2 else 0.5, p1=0.02 if i % 2 else 1.5,
                                     algo=C.TDX_ALGO_ICDF16 | flag))
        launches = run_descs(descs, bufs)
        assert launches == (2 if flag == 0 else 1)  # table kernel + direct kernel (small descriptor)
        outs[flag] = bufs
    for a, b in zip(outs[0], outs[C.TDX_ALGO_NOLUT]):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    # and a table-kernel shard boundary in the middle of a vector
    n = (1 << 27) + (1 << 21) + 3
    full = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
    run_descs([C.make_desc(full.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_count=n, seed=5, offset=8, p1=0.02)], [full])
    cut = (1 << 27) + 5
    a = torch.zeros(cut, dtype=TORCH_DT[dtype], device="cuda")
    b = torch.zeros(n - cut, dtype=TORCH_DT[dtype], device="cuda")
    run_descs([C.make_desc(a.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_count=cut, seed=5, offset=8, p1=0.02),
               C.make_desc(b.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_begin=cut, elem_count=n - cut, seed=5,
                           offset=8, p1=0.02)], [a, b])
    assert torch.equal(torch.cat([a, b]).view(torch.int16), full.view(torch.int16))


def test_table_kernel_preassigned_shares_and_folded_programs_match_the_direct_kernels():
    """A launch of >= 16 tiles (4 MiB) per SM takes the two-phase work list: every CTA walks its own
    pre-assigned share (broken where descriptors end), the tail comes from the work counter, and the
    module's constant fills and index programs ride in the same list.  Same bits as the direct
    kernels (TDX_ALGO_NOLUT: RNG, fill and iota kernels of their own), descriptor by descriptor --
    unaligned sizes, two parameter sets (table rebuilds inside a share), a shard that starts inside a
    vector."""
    dtype = C.TDX_BF16
    sizes = [(1 << 27) + 9, (1 << 26) + (1 << 25) + 12345, (1 << 27) + 7, (1 << 21) + 3, 3 << 24]
    assert sum(sizes) * 2 >= 148 * 16 * (256 << 10) * 1.2  # the pre-assigned path, with room to spare
    epi = [(C.TDX_EPI_MUL, 1.0 / 128), (C.TDX_EPI_RPOW, 500000.0), (C.TDX_EPI_RECIP, 0.0), (C.TDX_EPI_MUL, 1.0)]
    outs, counts = {}, {}
    for flag in (0, C.TDX_ALGO_NOLUT):
        bufs, descs = [], []
        for i, n in enumerate(sizes):
            t = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
            bufs.append(t)
            descs.append(C.make_desc(t.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_count=n, elem_begin=5 if i == 4 else 0,
                                     seed=91, offset=4096 * i, p0=0.0, p1=0.02 if i % 2 else 0.03,
                                     algo=C.TDX_ALGO_ICDF16 | flag))
        ones = torch.zeros(4099, dtype=torch.bfloat16, device="cuda")
        freq = torch.zeros(64, dtype=torch.float32, device="cuda")
        ids = torch.zeros(1000, dtype=torch.int64, device="cuda")
        bufs += [ones, freq, ids]
        descs += [C.make_desc(ones.data_ptr(), dtype=dtype, src=C.TDX_SRC_CONST, elem_count=ones.numel(), fill_bits=0x3F80, fill_itemsize=2),
                  C.make_desc(freq.data_ptr(), dtype=C.TDX_F32, src=C.TDX_SRC_IOTA, elem_count=64, p0=0, p1=2, epi=epi),
                  C.make_desc(ids.data_ptr(), dtype=C.TDX_I64, src=C.TDX_SRC_IOTA, elem_count=1000, p0=3, p1=-2)]
        counts[flag] = run_descs(descs, bufs)
        outs[flag] = bufs
    assert counts[0] == 1 and counts[C.TDX_ALGO_NOLUT] == 3  # one table launch carries everything
    for a, b in zip(outs[0], outs[C.TDX_ALGO_NOLUT]):
        assert torch.equal(a.view(torch.uint8), b.view(torch.uint8))
    assert torch.equal(outs[0][5], torch.ones(4099, dtype=torch.bfloat16, device="cuda"))
    assert torch.equal(outs[0][7], 3 - 2 * torch.arange(1000, device="cuda"))
    # launching the same plan twice gives the same bits (the work counter is put back; shares are static)
    again = [torch.zeros_like(t) for t in outs[0][:5]]
    descs2 = [C.make_desc(t.data_ptr(), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_count=n, elem_begin=5 if i == 4 else 0, seed=91,
                          offset=4096 * i, p0=0.0, p1=0.02 if i % 2 else 0.03, algo=C.TDX_ALGO_ICDF16)
              for i, (t, n) in enumerate(zip(again, sizes))]
    lib = C.load()
    ws = torch.empty(lib.tdx_init_workspace_bytes(len(descs2)), dtype=torch.uint8, device="cuda")
    plan = C.TdxPlan()
    stream = torch.cuda.current_stream().cuda_stream
    arr = (C.TdxInitDesc * len(descs2))(*descs2)
    C.check(lib.tdx_plan_upload(arr, len(descs2), ws.data_ptr(), ws.numel(), stream, ctypes.byref(plan)))
    for _ in range(2):
        for t in again:
            t.zero_()
        C.check(lib.tdx_plan_launch(ctypes.byref(plan), ws.data_ptr(), stream))
        torch.cuda.synchronize()
        for a, b in zip(again, outs[0][:5]):
            assert torch.equal(a.view(torch.int16), b.view(torch.int16))


def test_table_kernel_with_thousands_of_small_descriptors_in_one_launch():
    """3,600 table-eligible descriptors of ~0.5 MiB (two tiles and a ragged end each, so ~3 work-list
    entries per descriptor and ~70 per CTA: every pre-assigned share crosses dozens of descriptor
    boundaries); three parameter sets make tables change inside a share.  Same bits as the direct
    kernel, nothing written outside the descriptors."""
    dtype, n_desc = C.TDX_BF16, 3600
    sizes = [(1 << 18) + 8 * (i % 7) + (i % 3) for i in range(n_desc)]
    offs = np.concatenate([[0], np.cumsum([(n + 7) // 8 * 8 for n in sizes])])  # 16-byte aligned starts
    outs = {}
    for flag in (0, C.TDX_ALGO_NOLUT):
        buf = torch.full((int(offs[-1]),), -1, dtype=torch.int16, device="cuda").view(torch.bfloat16)  # NaN pattern
        descs = [C.make_desc(buf.data_ptr() + 2 * int(offs[i]), dtype=dtype, src=C.TDX_SRC_NORMAL, elem_count=n, seed=123,
                             offset=64 * i, p0=0.0, p1=(0.02, 0.03, 0.5)[(i // 40) % 3], algo=C.TDX_ALGO_ICDF16 | flag)
                 for i, n in enumerate(sizes)]
        assert run_descs(descs, [buf]) == 1
        outs[flag] = buf
    assert torch.equal(outs[0].view(torch.int16), outs[C.TDX_ALGO_NOLUT].view(torch.int16))
    # every element of every descriptor was written (finite), the padding between descriptors was not
    inside = torch.ones(int(offs[-1]), dtype=torch.bool, device="cuda")
    for i, n in enumerate(sizes):
        inside[int(offs[i]) + n:int(offs[i + 1])] = False
    bits = outs[0].view(torch.int16)
    assert bool((bits[inside] != -1).all()) and bool((bits[~inside] == -1).all())
    assert bool(torch.isfinite(outs[0][inside].float()).all())


@pytest.mark.parametrize("dtype,src", [(C.TDX_BF16, C.TDX_SRC_NORMAL), (C.TDX_F32, C.TDX_SRC_UNIFORM)])
def test_beyond_2_to_32_elements(dtype, src):
    """BASELINE config #5 reaches 16 GB tensors: global element indices above 2^32 must index the
    Philox stream with 64 bits.  A shard that starts just below the 2^32 boundary of a (virtual)
    2^33-element tensor is compared with the oracle element by element, and with the same region
    produced inside a larger launch."""
    base = (1 << 32) - 1000  # global index of the first element this launch writes
    n = (1 << 21) + 7        # crosses 2^32 (and the table kernel's size threshold for bf16)
    buf = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
    d = C.make_desc(buf.data_ptr(), dtype=dtype, src=src, elem_begin=base, elem_count=n, seed=3, offset=16,
                    p0=-1.0 if src == C.TDX_SRC_UNIFORM else 0.0, p1=1.0)
    run_descs([d], [buf])
    got = gpu_bits(buf, dtype)
    for lo, cnt in ((0, 4096), (990, 64), (n - 4096, 4096)):  # around the boundary and at both ends
        piece = C.make_desc(0, dtype=dtype, src=src, elem_begin=base + lo, elem_count=cnt, seed=3, offset=16,
                            p0=-1.0 if src == C.TDX_SRC_UNIFORM else 0.0, p1=1.0)
        diff = got[lo:lo + cnt].astype(np.int64) - O.generate(piece).astype(np.int64)
        assert np.abs(diff).max() <= (0 if src == C.TDX_SRC_UNIFORM else 1)
    # the stream is not periodic in 2^32: the same offsets 2^32 elements earlier differ
    early = torch.zeros(4096, dtype=TORCH_DT[dtype], device="cuda")
    run_descs([C.make_desc(early.data_ptr(), dtype=dtype, src=src, elem_begin=base - (1 << 32) + 1000, elem_count=4096,
                           seed=3, offset=16, p0=-1.0 if src == C.TDX_SRC_UNIFORM else 0.0, p1=1.0)], [early])
    assert (gpu_bits(early, dtype) != got[1000:1000 + 4096]).mean() > 0.9


def test_sixteen_gigabyte_tensor_statistics():
    """The largest sweep point: one 16 GiB bf16 normal tensor (8.6 G elements), written by one launch."""
    n = 1 << 33
    free, _ = torch.cuda.mem_get_info()
    if free < (n * 2) + (4 << 30):
        pytest.skip("not enough free HBM")
    buf = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    d = C.make_desc(buf.data_ptr(), dtype=C.TDX_BF16, src=C.TDX_SRC_NORMAL, elem_count=n, seed=11, offset=0, p0=0.0, p1=0.02)
    run_descs([d], [buf])
    for lo in (0, (1 << 32) - (1 << 20), n - (1 << 22)):
        x = buf[lo:lo + (1 << 22)].float()
        assert abs(x.mean().item()) < 5 * 0.02 / 2048 and abs(x.std().item() / 0.02 - 1) < 5e-3
    tail = C.make_desc(0, dtype=C.TDX_BF16, src=C.TDX_SRC_NORMAL, elem_begin=n - 4096, elem_count=4096, seed=11,
                       offset=0, p0=0.0, p1=0.02)
    diff = gpu_bits(buf[n - 4096:], C.TDX_BF16).astype(np.int64) - O.generate(tail).astype(np.int64)
    assert np.abs(diff).max() <= 1
    del buf
    torch.cuda.empty_cache()


@pytest.mark.parametrize("dtype", [C.TDX_BF16, C.TDX_F16])
@pytest.mark.parametrize("src", [C.TDX_SRC_UNIFORM, C.TDX_SRC_NORMAL])
def test_wide_16bit_outputs_are_the_rounded_fp32_tensor(dtype, src):
    """TDX_ALGO_WIDE32: a 16-bit tensor generated with the fp32 stream and arithmetic equals the
    fp32 tensor of the same (seed, offset) cast to the dtype -- with and without fp32 epilogue
    steps -- and matches the oracle."""
    n = 100003
    kw = dict(elem_count=n, seed=21, offset=64, p0=-0.25 if src == C.TDX_SRC_UNIFORM else 0.5, p1=0.75)
    for epi32, epi16 in (((), ()), (((C.TDX_EPI_MUL, 0.02), (C.TDX_EPI_ADD, 1.0)),
                                  ((C.TDX_EPI_MUL | C.TDX_EPI_NOROUND, 0.02), (C.TDX_EPI_ADD | C.TDX_EPI_NOROUND, 1.0)))):
        f32 = torch.zeros(n, dtype=torch.float32, device="cuda")
        t16 = torch.zeros(n, dtype=TORCH_DT[dtype], device="cuda")
        d32 = C.make_desc(f32.data_ptr(), dtype=C.TDX_F32, src=src, epi=epi32, **kw)
        d16 = C.make_desc(t16.data_ptr(), dtype=dtype, src=src, algo=C.TDX_ALGO_WIDE32, epi=epi16,
                          flags=C.TDX_FLAG_SRC_NOROUND, **kw)
        run_descs([d32, d16], [f32, t16])
        assert torch.equal(t16, f32.to(TORCH_DT[dtype]))
        exp = O.generate(d16)
        diff = gpu_bits(t16, dtype).astype(np.int64) - exp.astype(np.int64)
        if src == C.TDX_SRC_UNIFORM:
            assert not diff.any()
        else:
            # MUFU vs libm: an ABSOLUTE error ~1e-6 std on the fp32 value (see the fp32 normal test);
            # near x = 0 that is many ulps of the 16-bit result, so compare values, not ulps
            g, e = as_float(gpu_bits(t16, dtype), dtype), as_float(exp, dtype)
            ulp = np.abs(e) * (2.0 ** -7 if dtype == C.TDX_BF16 else 2.0 ** -10)  # spacing is at most this
            assert np.all(np.abs(g - e) <= ulp + 1e-5 * kw["p1"] * (1 + np.abs(e)))
            assert (diff != 0).mean() < 0.002
