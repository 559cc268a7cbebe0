"""SURVEY 8f.1 -- FSDP hand-off without the intermediate copy (needs >= 2 GPUs: run with
`gpurun --gpus 2`; skipped on a single-GPU box).

Flow: build the model on the meta device, `fully_shard` it (parameters become dim-0 sharded
DTensors), `to_empty` on the GPU, then every rank fills the local shards FSDP owns IN PLACE from an
InitPlan.  The gathered parameters must equal the unsharded `materialize_module` result bit for
bit, and nobody ever allocated a full tensor."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, plan_path, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from torch.distributed.fsdp import fully_shard

    from oracle import cases
    from torchdistx_b200 import parallel
    from torchdistx_b200.plan import InitPlan, init_sharded_module

    with torch.device("meta"):
        model = cases.build("mlp_stack", "fp32")
    for layer in model:
        if any(True for _ in layer.parameters()):
            fully_shard(layer)
    fully_shard(model)
    model.to_empty(device=torch.device("cuda", rank))
    torch.manual_seed(1000 + rank)          # ranks disagree ...
    parallel.sync_rng(torch.device("cuda", rank))  # ... until the 16-byte broadcast
    torch.manual_seed(77)
    mem0 = torch.cuda.max_memory_allocated()
    init_sharded_module(model, InitPlan.load(plan_path), rank, world, device=torch.device("cuda", rank))
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - mem0
    full = {n: p.full_tensor().cpu() for n, p in model.named_parameters()}
    if rank == 0:
        torch.save({"full": full, "peak": peak}, os.path.join(outdir, "gathered.pt"))
    # the sharded module trains
    y = model(torch.randn(4, 64, device="cuda"))
    y.sum().backward()
    dist.barrier()
    dist.destroy_process_group()


def test_fsdp2_local_shards_are_initialised_in_place(tmp_path):
    from oracle import cases
    from torchdistx_b200.deferred_init import deferred_init, materialize_module
    from torchdistx_b200.plan import InitPlan

    plan_path = str(tmp_path / "plan.json")
    InitPlan.from_module(deferred_init(lambda: cases.build("mlp_stack", "fp32", "cuda"))).save(plan_path)
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), plan_path, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(tmp_path / "gathered.pt")

    m = deferred_init(lambda: cases.build("mlp_stack", "fp32", "cuda"))
    torch.manual_seed(77)
    materialize_module(m)
    for n, p in m.named_parameters():
        assert torch.equal(got["full"][n], p.detach().cpu()), n
    total = sum(p.numel() * p.element_size() for p in m.parameters())
    assert got["peak"] < 0.25 * total + (1 << 20)  # only the descriptor workspace was allocated
