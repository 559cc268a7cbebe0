"""The reference's actual caller, unmodified: PyTorch's FSDP1 probes `import torchdistx`
($TORCH/distributed/fsdp/_init_utils.py:53-57), asks `fake.is_fake(param)` (:574-594) and calls
`deferred_init.materialize_module(module, check_fn=...)` (:881-885) before it flattens and shards.
Two gloo ranks on the CPU wrap a DEFERRED module with FullyShardedDataParallel and nothing else: the
repo's `torchdistx` shim is what FSDP finds.

Checked: FSDP materialised the module through this engine (call counters), the unsharded parameters
are the eager same-seed model bit for bit (CPU tensors: T0'), each rank's flat-parameter shard is
`FlatParamHandle._get_shard` of the flattened eager model, nested wrapping drives `check_fn` (inner
FSDP units are materialised by their own wrapper, the outer call must skip them), and a forward /
backward / optimizer step runs."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class Block(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.fc1 = nn.Linear(d, 2 * d)
        # a dead kaiming pass + a live normal (the HF idiom).  No bias: materialisation replays a tensor's
        # whole program when its turn comes (reference deferred_init.cc:541-622), so a re-initialisation
        # that an eager constructor would run AFTER another tensor's draw would order the stream differently
        self.fc2 = nn.Linear(2 * d, d, bias=False)
        nn.init.normal_(self.fc2.weight, 0.0, 0.02)
        self.norm = nn.LayerNorm(d)

    def forward(self, x):
        return self.norm(x + self.fc2(torch.relu(self.fc1(x))))


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.embed = nn.Embedding(50, 24, padding_idx=0)
        self.blocks = nn.ModuleList([Block(24) for _ in range(3)])
        self.head = nn.Linear(24, 50, bias=False)
        self.register_buffer("scale", torch.full((1,), 0.5))

    def forward(self, ids):
        x = self.embed(ids) * self.scale
        for b in self.blocks:
            x = b(x)
        return self.head(x)


def _worker(rank, world, port, nested):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.distributed.fsdp._init_utils as init_utils
    from torch.distributed.fsdp import FullyShardedDataParallel as FSDP
    from torch.distributed.fsdp._flat_param import FlatParamHandle
    from torch.distributed.fsdp.wrap import ModuleWrapPolicy

    import torchdistx_b200.deferred_init as engine
    from torchdistx.deferred_init import deferred_init, is_deferred
    from torchdistx.fake import is_fake

    # FSDP found the shim, and the shim is this engine
    assert init_utils._TORCHDISTX_AVAIL and init_utils.deferred_init.materialize_module is engine.materialize_module

    calls = []
    real = engine._C.materialize_module

    def counting(module, buffers_only, check_fn, *rest):
        calls.append((type(module).__name__, check_fn is not None))
        return real(module, buffers_only, check_fn, *rest)

    engine._C.materialize_module = counting
    try:
        model = deferred_init(Net)
        assert is_deferred(model) and all(is_fake(p) for p in model.parameters())
        torch.manual_seed(1234)  # same seed on every rank: FSDP1 materialises the full module per rank, then shards
        kw = dict(auto_wrap_policy=ModuleWrapPolicy({Block})) if nested else {}
        fsdp = FSDP(model, device_id=torch.device("cpu"), use_orig_params=False, **kw)
    finally:
        engine._C.materialize_module = real
    assert not is_deferred(fsdp)
    # one engine call per FSDP unit, each with FSDP's check_fn (skip submodules that are FSDP units already)
    assert calls == ([("Block", True)] * 3 if nested else []) + [("Net", True)], calls

    # eager twin: same seed, same order of materialisation (auto-wrap initialises the blocks first,
    # post-order, then the root's own tensors)
    torch.manual_seed(1234)
    if nested:
        # replay the order FSDP used: blocks 0..2 (children before the root unit), then embed, head
        # -- a module built eagerly draws in CONSTRUCTION order instead, so build the pieces in that order
        eager = Net.__new__(Net)
        nn.Module.__init__(eager)
        blocks = [Block(24) for _ in range(3)]
        eager.embed = nn.Embedding(50, 24, padding_idx=0)
        eager.blocks = nn.ModuleList(blocks)
        eager.head = nn.Linear(24, 50, bias=False)
        eager.register_buffer("scale", torch.full((1,), 0.5))
    else:
        eager = Net()
    with FSDP.summon_full_params(fsdp):
        got = dict(fsdp.module.named_parameters())
        for name, p in eager.named_parameters():
            cand = [k for k in got if k.replace("_fsdp_wrapped_module.", "") == name]
            assert len(cand) == 1, (name, list(got))
            assert torch.equal(got[cand[0]].detach(), p.detach()), name
        assert torch.equal(dict(fsdp.module.named_buffers())["scale"], eager.scale)
        assert bool((fsdp.module.embed.weight[0] == 0).all())  # the padding row went through the engine too

    # each rank's shard of the root flat parameter is torch's own _get_shard of the flattened eager tensors
    handle = fsdp._handle
    root_params = [p for n, p in eager.named_parameters() if not (nested and n.startswith("blocks."))]
    flat = torch.cat([p.detach().reshape(-1) for p in root_params])
    exp, padded = FlatParamHandle._get_shard(flat, rank, world)
    assert torch.equal(handle.flat_param._local_shard, exp), (rank, padded)

    # and it trains
    opt = torch.optim.SGD(fsdp.parameters(), lr=0.1)
    before = handle.flat_param._local_shard.clone()
    loss = fsdp(torch.randint(1, 50, (4, 7))).float().pow(2).mean()
    loss.backward()
    opt.step()
    assert torch.isfinite(loss) and not torch.equal(handle.flat_param._local_shard, before)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nested", [False, True], ids=["one_unit", "auto_wrapped_blocks"])
def test_fsdp1_materialises_a_deferred_module_through_the_shim(nested):
    mp.spawn(_worker, args=(2, _free_port(), nested), nprocs=2, join=True)


def test_flat_shard_layout_is_torchs_own():
    """`materialize_flat_shard` is held (tests/test_engine_round2_gpu.py, on the GPU) to this layout:
    parameters flattened in order, each start aligned to `align_numel`, the flat vector cut into
    `world` chunks of ceil(total / world) elements, the tail padded.  Here the same arithmetic against
    PyTorch's own `FlatParamHandle.flatten_tensors` + `_get_shard`
    ($TORCH/distributed/fsdp/_flat_param.py): same offsets, same chunk size, every parameter element in
    the same place of the same rank's shard.  One difference, in elements nobody reads: PyTorch fills
    ALIGNMENT padding with its debug value 42 (`_FLAT_PARAM_PADDING_VALUE`), this engine with zeros
    (the tail padding of the unaligned layout is zeros in both)."""
    import types

    from torch.distributed.fsdp._flat_param import FlatParamHandle

    def torch_shards(tensors, world, align):
        dummy = types.SimpleNamespace(world_size=world, _use_orig_params=True)
        dummy._validate_tensors_to_flatten = types.MethodType(FlatParamHandle._validate_tensors_to_flatten, dummy)
        flat = FlatParamHandle.flatten_tensors(dummy, tensors, align)
        return [FlatParamHandle._get_shard(flat, r, world)[0] for r in range(world)]

    def engine_layout(tensors, world, align):  # (the restatement of the GPU test, plus a mask of real elements)
        parts, mask, offsets, total = [], [], [], 0
        for p in tensors:
            if align > 1 and total % align:
                pad = align - total % align
                parts.append(torch.zeros(pad, dtype=p.dtype))
                mask.append(torch.zeros(pad, dtype=torch.bool))
                total += pad
            offsets.append(total)
            parts.append(p.detach().flatten())
            mask.append(torch.ones(p.numel(), dtype=torch.bool))
            total += p.numel()
        chunk = -(-total // world)
        flat = torch.cat(parts + [torch.zeros(chunk * world - total, dtype=parts[0].dtype)])
        real = torch.cat(mask + [torch.zeros(chunk * world - total, dtype=torch.bool)])
        return [flat[r * chunk:(r + 1) * chunk] for r in range(world)], [real[r * chunk:(r + 1) * chunk] for r in range(world)]

    torch.manual_seed(0)
    tensors = [torch.randn(7, 3), torch.randn(5), torch.randn(4, 4), torch.randn(1), torch.randn(2, 3, 2)]
    for world in (1, 2, 3, 8):
        for align in (0, 4, 8):
            theirs = torch_shards(tensors, world, align)
            mine, real = engine_layout(tensors, world, align)
            for r in range(world):
                assert theirs[r].shape == mine[r].shape, (world, align, r)
                assert torch.equal(theirs[r][real[r]], mine[r][real[r]]), (world, align, r)
                pad = theirs[r][~real[r]]
                assert bool(((pad == 42) | (pad == 0)).all()) and bool((mine[r][~real[r]] == 0).all())
                if align == 0:
                    assert torch.equal(theirs[r], mine[r])
