"""Multi-GPU plumbing for sharded materialisation.

The path shards naturally: element g of a tensor is a pure function of
(seed, Philox offset of its RNG op, g), so every rank can build its own dim-0 chunk with no
data exchange.  The only thing ranks must agree on is the generator state the offsets are
drawn from; `sync_rng` is that one collective (16 bytes, broadcast from rank 0 over NCCL/NVLink,
or gloo on CPU test rigs).  The reference has no multi-GPU materialise path at all
(SURVEY.md section 8e); FSDP calls it per rank and shards afterwards
($TORCH/distributed/fsdp/_init_utils.py:574-607).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def _generator(device: torch.device) -> torch.Generator:
    if device.type == "cuda":
        idx = device.index if device.index is not None else torch.cuda.current_device()
        return torch.cuda.default_generators[idx]
    return torch.default_generator


def rng_state(device) -> Tuple[int, int]:
    """(seed, offset) of the default generator the engine draws Philox offsets from."""
    device = torch.device(device)
    g = _generator(device)
    return int(g.initial_seed()), int(g.get_offset()) if device.type == "cuda" else 0


# (device, group) -> (seed, offset) the ranks agreed on at the last broadcast
_agreed: Dict[Tuple[str, int], Tuple[int, int]] = {}


def sync_rng(device, group: Optional[dist.ProcessGroup] = None, src: int = 0, *, force: bool = False) -> Tuple[int, int]:
    """Makes every rank of `group` hold rank `src`'s (seed, offset).  Returns this rank's pair.

    Call before `materialize_module(..., shard=(rank, world))`.  With the same state and the same
    module, all ranks derive identical per-tensor offsets from the traversal order, so the shards
    they write are exactly the slices of one (never materialised) unsharded tensor.

    The first call for a (device, group) is the path's one collective: a 16-byte broadcast.  After
    it the ranks' generators advance in lock step -- every RNG pass consumes a function of the
    GLOBAL element count only (planner.cc assign_rng), whatever a rank's share of the rows is -- so
    later calls find the state they left behind (same seed, offset not behind the agreed one) and
    return without communicating or synchronising the device; the broadcast is repeated when the
    generator was re-seeded or rewound in between, or with ``force=True``.  `check_agreement`
    verifies that the ranks still agree (one small all-reduce; use it in tests and after a timed
    region, not inside one).
    """
    device = torch.device(device)
    seed, offset = rng_state(device)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    key = (str(device), id(group))
    if multi and not force:
        agreed = _agreed.get(key)
        if agreed is not None and agreed[0] == seed and offset >= agreed[1]:
            return seed, offset
    if multi:
        # int64 transport: seeds are unsigned 64-bit in torch; fold to two's complement
        to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v
        backend = dist.get_backend(group)
        buf_device = device if backend == "nccl" else torch.device("cpu")
        buf = torch.tensor([to_i64(seed), to_i64(offset)], dtype=torch.int64, device=buf_device)
        dist.broadcast(buf, src=src, group=group)
        seed, offset = (int(v) & ((1 << 64) - 1) for v in buf.tolist())
        _agreed[key] = (seed, offset)
    g = _generator(device)
    g.manual_seed(seed)
    if device.type == "cuda":
        g.set_offset(offset)
    return seed, offset


def check_agreement(device, group: Optional[dist.ProcessGroup] = None) -> bool:
    """All ranks hold the same (seed, offset)?  (min == max over the group)"""
    seed, offset = rng_state(device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return True
    backend = dist.get_backend(group)
    buf_device = torch.device(device) if backend == "nccl" else torch.device("cpu")
    v = torch.tensor([seed & ((1 << 62) - 1), offset & ((1 << 62) - 1)], dtype=torch.int64, device=buf_device)
    lo, hi = v.clone(), v.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool(torch.equal(lo, hi))
