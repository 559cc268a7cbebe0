"""N > 1 host logic on CPU: two gloo ranks agree on the generator state (the one collective of
the path) and build disjoint dim-0 chunks that concatenate to the unsharded module.  On CPU the
tensors come from generic ATen replay and are chunked afterwards; the geometry, traversal order,
buffer replication and RNG agreement logic are the same as on the GPU path."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def sd_shapes(key, case):
    from oracle import cases

    return dict(cases.build(case, "fp32").named_parameters())[key].shape


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cases
    from torchdistx_b200 import parallel
    from torchdistx_b200.deferred_init import deferred_init, is_deferred, materialize_module

    torch.manual_seed(100 + rank)  # ranks start with DIFFERENT seeds ...
    assert not parallel.check_agreement("cpu") or world == 1
    seed, offset = parallel.sync_rng("cpu")  # ... and agree after the broadcast
    assert parallel.check_agreement("cpu") and seed == 100

    m = deferred_init(lambda: cases.build("init_zoo", "fp32"))
    materialize_module(m, shard=(rank, world))
    assert not is_deferred(m)
    sd = {k: v.detach().clone() for k, v in list(m.named_parameters()) + list(m.named_buffers())}
    sd["__is_param__"] = sorted(k for k, _ in m.named_parameters())
    torch.save(sd, os.path.join(outdir, f"rank{rank}.pt"))

    # the same through a DeviceMesh (FSDP2 / DTensor Shard(0) layout), wrapped as DTensors: the mesh's
    # group carries the seed agreement, full_tensor() gathers what the ranks built
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import DTensor

    mesh = init_device_mesh("cpu", (world,))
    torch.manual_seed(200 + rank)  # disagree again: materialize_module(device_mesh=...) syncs by itself
    m2 = deferred_init(lambda: cases.build("mlp_stack", "fp32"))
    materialize_module(m2, device_mesh=mesh, as_dtensor=True)
    assert parallel.check_agreement("cpu")
    assert parallel.sync_rng("cpu") == parallel.rng_state("cpu")  # derived locally: no second broadcast needed
    gathered = {}
    for k, p in m2.named_parameters():
        assert isinstance(p, torch.nn.Parameter) and isinstance(p.data, DTensor) and p.shape == sd_shapes(k, "mlp_stack"), k
        gathered[k] = p.full_tensor().detach().clone()
    if rank == 0:
        torch.save(gathered, os.path.join(outdir, "mesh_gathered.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_gloo_ranks_build_complementary_shards(world, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    shards = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]

    from oracle import cases
    from torchdistx_b200.deferred_init import deferred_init, materialize_module

    torch.manual_seed(100)
    full = deferred_init(lambda: cases.build("init_zoo", "fp32"))
    materialize_module(full)
    params = set(shards[0].pop("__is_param__"))
    for s in shards[1:]:
        s.pop("__is_param__")
    for k, t in list(full.named_parameters()) + list(full.named_buffers()):
        t = t.detach()
        if k in params and t.dim() > 0:
            chunks = torch.chunk(t, world, 0)
            for r in range(world):
                exp = chunks[r] if r < len(chunks) else t[:0]
                assert torch.equal(shards[r][k], exp), (k, r)
        else:
            for r in range(world):
                assert torch.equal(shards[r][k], t), (k, r)

    # DeviceMesh run: rank 0's seed (200) wins; the gathered DTensors are the unsharded module
    torch.manual_seed(200)
    full2 = deferred_init(lambda: cases.build("mlp_stack", "fp32"))
    materialize_module(full2)
    gathered = torch.load(tmp_path / "mesh_gathered.pt")
    for k, t in full2.named_parameters():
        assert torch.equal(gathered[k], t.detach()), k
