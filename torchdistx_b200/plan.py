"""Serialisable initialisation plans.

The reference's recording cannot leave the process: every op holds a ``std::function`` over the
dispatcher and an ``OperatorHandle&`` (reference src/cc/torchdistx/deferred_init.cc:159, 217-225),
so re-materialising a model -- elastic restart, a new rank joining, another seed -- means
re-running the Python constructor under ``deferred_init``.  Here the planner's verdict per tensor
is plain data (source, parameters, epilogue, RNG passes on the chain), so it can be written to
disk and replayed later with nothing but the C ABI:

    plan = InitPlan.from_module(deferred_model)        # works without a GPU
    plan.save("llama3-8b.init.json")
    ...
    tensors = InitPlan.load("llama3-8b.init.json").materialize(device="cuda", shard=(rank, world))

Under the same generator state ``InitPlan.materialize`` produces exactly the bits
``materialize_module`` produces (tests/test_plan_gpu.py): same traversal order, same Philox
offset bookkeeping, same kernels.  Tensors whose program is not fusible (tiny ``arange``-style
buffers such as rotary ``inv_freq``) are evaluated once when the plan is built and stored by
value (refused above ``max_embedded_bytes``).  Building a plan therefore MATERIALISES those few
tensors on the recording (on their recorded device); it does so under a forked RNG state, so the
caller's generators are left as they were, and it refuses programs that draw random numbers
(embedding one sample would freeze it).
"""
from __future__ import annotations

import base64
import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
from torch.nn import Module

from . import _C, _cabi

_DTYPES = {
    "Float": torch.float32, "BFloat16": torch.bfloat16, "Half": torch.float16, "Double": torch.float64,
    "Long": torch.int64, "Int": torch.int32, "Short": torch.int16, "Char": torch.int8,
    "Byte": torch.uint8, "Bool": torch.bool,
}
_TDX_DTYPE = {torch.float32: _cabi.TDX_F32, torch.bfloat16: _cabi.TDX_BF16, torch.float16: _cabi.TDX_F16}
_RAW = {1: _cabi.TDX_RAW8, 2: _cabi.TDX_RAW16, 4: _cabi.TDX_RAW32, 8: _cabi.TDX_RAW64}
_DTYPE_NAMES = {v: k for k, v in _DTYPES.items()}
FORMAT = "torchdistx_b200.InitPlan/2"


def offset_increment(numel: int) -> int:
    """Generator offset one RNG pass over `numel` elements consumes (planner.cc assign_rng)."""
    blocks = (numel + 3) // 4
    return ((blocks + 3) // 4) * 4 + 4


@dataclass
class PlanEntry:
    name: str
    kind: str  # "param" | "buffer"
    sizes: List[int]
    dtype: str
    source: str  # "const" | "uniform" | "normal" | "uninit" | "value"
    p0: float = 0.0
    p1: float = 0.0
    epilogue: List[Tuple[int, float, float]] = field(default_factory=list)
    const_bytes: str = ""  # base64, one element
    value: str = ""  # base64 of the whole tensor for "value" entries
    rng_numels: List[int] = field(default_factory=list)
    requires_grad: bool = False
    alias_of: Optional[str] = None  # tied parameters: same tensor as an earlier entry
    wide: bool = False  # fp32 source cast to a 16-bit dtype (TDX_ALGO_WIDE32)
    src_noround: bool = False  # TDX_FLAG_SRC_NOROUND
    # Identity of every RNG pass of `rng_numels` inside the plan: a deepcopy / clone shares its
    # source's passes (the copy is bit-identical and consumes nothing of the generator).
    rng_ids: List[int] = field(default_factory=list)
    # The tensor as disjoint element ranges, each with its own source (`w.normal_();
    # w[padding_idx].zero_()` is three); empty = one segment described by the fields above.
    # {"begin", "end", "origin", "source", "p0", "p1", "epilogue", "const_bytes", "wide",
    #  "src_noround", "rng_pass" (index into rng_numels, -1: none)}
    segments: List[dict] = field(default_factory=list)

    def segment_list(self) -> List[dict]:
        if self.segments:
            return self.segments
        numel = 1
        for d in self.sizes:
            numel *= d
        return [dict(begin=0, end=numel, origin=0, source=self.source, p0=self.p0, p1=self.p1,
                     epilogue=self.epilogue, const_bytes=self.const_bytes, wide=self.wide,
                     src_noround=self.src_noround, rng_pass=len(self.rng_numels) - 1)]


def shard_range(e: "PlanEntry", shard: Optional[Tuple[int, int]]) -> Tuple[int, int, List[int]]:
    """(first element, element count, sizes) of this rank's part of an entry: parameters are cut on
    dim 0 like ``torch.chunk`` (ceil(d0 / world) rows per rank, trailing ranks may get none), buffers
    are replicated -- the layout of ``materialize_module(shard=...)`` and of FSDP2's ``Shard(0)``."""
    sizes = list(e.sizes)
    numel = 1
    for s in sizes:
        numel *= s
    if shard is None or shard[1] <= 1 or e.kind != "param" or not sizes:
        return 0, numel, sizes
    rank, world = shard
    d0 = sizes[0]
    inner = numel // d0 if d0 else 0
    per = -(-d0 // world)
    start = min(d0, rank * per)
    rows = min(per, d0 - start)
    return start * inner, rows * inner, [rows] + sizes[1:]


def value_tensor(e: "PlanEntry", shard: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """A by-value entry as a CPU tensor: whole, or this rank's dim-0 chunk of a parameter (what
    ``materialize_module(shard=...)`` does with a replayed parameter: narrow + clone)."""
    raw = torch.frombuffer(bytearray(base64.b64decode(e.value)), dtype=torch.uint8)
    t = raw.view(_DTYPES[e.dtype]).reshape(e.sizes)
    begin, count, sizes = shard_range(e, shard)
    if count != t.numel():
        t = t.reshape(-1)[begin: begin + count].reshape(sizes)
    return t


def assign_pass_offsets(e: "PlanEntry", assigned: Dict[int, int], offset: int) -> Tuple[List[int], int]:
    """Philox offsets of the entry's RNG passes, and the generator offset after them.  Every pass on
    the chain takes its slice of the stream once per pass identity: a clone names its source's
    passes again (same stream, nothing consumed)."""
    ids = e.rng_ids if e.rng_ids else [None] * len(e.rng_numels)
    pass_offset = []
    for pid, n in zip(ids, e.rng_numels):
        if pid is not None and pid in assigned:
            pass_offset.append(assigned[pid])
            continue
        pass_offset.append(offset)
        if pid is not None:
            assigned[pid] = offset
        offset += offset_increment(n)
    return pass_offset, offset


def entry_descriptors(e: "PlanEntry", base: int, begin: int, count: int, seed: int,
                      pass_offset: List[int]) -> list:
    """The C-ABI descriptors that write elements [begin, begin + count) of an entry into a buffer
    whose first byte is at address `base` (plain data in, ``TdxInitDesc`` out: no device needed)."""
    dtype = _DTYPES[e.dtype]
    isz = torch.empty((), dtype=dtype).element_size()
    descs = []
    for g in e.segment_list():
        lo, hi = max(g["begin"], begin), min(g["end"], begin + count)
        if lo >= hi or g["source"] == "uninit":
            continue
        dst = base + (lo - begin) * isz
        if g["source"] == "const":
            descs.append(_cabi.make_desc(
                dst, dtype=_RAW[isz], src=_cabi.TDX_SRC_CONST, elem_count=hi - lo,
                fill_bits=int.from_bytes(base64.b64decode(g["const_bytes"]), "little"), fill_itemsize=isz))
        elif g["source"] == "iota":  # arange and the index programs built on it (rotary inv_freq)
            descs.append(_cabi.make_desc(
                dst, dtype=_cabi.TDX_I64 if dtype == torch.int64 else _TDX_DTYPE[dtype],
                src=_cabi.TDX_SRC_IOTA, elem_begin=lo - g["origin"], elem_count=hi - lo,
                p0=g["p0"], p1=g["p1"], epi=g["epilogue"]))
        else:
            descs.append(_cabi.make_desc(
                dst, dtype=_TDX_DTYPE[dtype],
                src=_cabi.TDX_SRC_UNIFORM if g["source"] == "uniform" else _cabi.TDX_SRC_NORMAL,
                elem_begin=lo - g["origin"], elem_count=hi - lo, seed=seed,
                offset=pass_offset[g["rng_pass"]], p0=g["p0"], p1=g["p1"],
                epi=g["epilogue"], algo=_cabi.TDX_ALGO_WIDE32 if g["wide"] else 0,
                flags=_cabi.TDX_FLAG_SRC_NOROUND if g["src_noround"] else 0))
    return descs


class InitPlan:
    def __init__(self, entries: List[PlanEntry]):
        self.entries = entries

    # ------------------------------------------------------------------------------------- build
    @classmethod
    def from_module(cls, module: Module, max_embedded_bytes: int = 1 << 20) -> "InitPlan":
        """Builds the plan of a module created by ``deferred_init``.  Traversal order is the one
        ``materialize_module`` uses (children first, then own parameters, then own buffers)."""
        from .deferred_init import materialize_tensor

        entries: List[PlanEntry] = []
        seen: Dict[int, str] = {}

        def visit(mod: Module, prefix: str) -> None:
            for cname, child in mod._modules.items():
                if child is not None:
                    visit(child, f"{prefix}{cname}.")
            for kind, group in (("param", mod._parameters), ("buffer", mod._buffers)):
                for key, t in group.items():
                    if t is None:
                        continue
                    name = prefix + key
                    if id(t) in seen:
                        entries.append(PlanEntry(name, kind, list(t.shape), _DTYPE_NAMES[t.dtype], "alias", alias_of=seen[id(t)]))
                        continue
                    seen[id(t)] = name
                    info = dict(_C.plan_info(t))
                    if info["source"] in ("const", "uniform", "normal", "uninit", "iota") and info["fusible"]:
                        segs = [dict(begin=g["begin"], end=g["end"], origin=g["origin"], source=g["source"],
                                     p0=g["p0"], p1=g["p1"], epilogue=[tuple(e) for e in g["epilogue"]],
                                     const_bytes=base64.b64encode(g["const_bytes"]).decode(),
                                     wide=bool(g["wide"]), src_noround=bool(g["src_noround"]),
                                     rng_pass=int(g["rng_pass"])) for g in info["segments"]]
                        entries.append(PlanEntry(
                            name, kind, list(info["sizes"]), info["dtype"], info["source"], info["p0"], info["p1"],
                            [tuple(e) for e in info["epilogue"]],
                            base64.b64encode(info["const_bytes"]).decode(), "", list(info["rng_numels"]),
                            bool(info["requires_grad"]), None, bool(info["wide"]), bool(info["src_noround"]),
                            [int(i) for i in info["rng_op_ids"]], segs if len(segs) > 1 else []))
                        continue
                    # not fusible (or already real): evaluate now and store by value -- under a forked
                    # RNG state, and never a program that draws random numbers (one frozen sample)
                    if info["deferred"]:
                        if any(n.startswith(("aten::rand", "aten::normal", "aten::uniform", "aten::bernoulli",
                                             "aten::multinomial", "aten::dropout"))
                               for n in (h.split(" ")[0] for h in _C.storage_history(t))):
                            raise ValueError(
                                f"'{name}' has a random initialisation program the planner cannot fold "
                                f"(first unfusable op: {info['first_unfusable_op'] or 'n/a'}); a plan cannot embed it")
                        devs = [torch.device(info["device"])] if info["device"].startswith("cuda") and torch.cuda.is_available() else []
                        with torch.random.fork_rng(devices=devs):
                            before = [torch.get_rng_state()] + [torch.cuda.get_rng_state(d) for d in devs]
                            real = materialize_tensor(t)
                            after = [torch.get_rng_state()] + [torch.cuda.get_rng_state(d) for d in devs]
                        # (the names above only see this tensor's own storage: `randn(n).sub_(c) * 2 + 3` draws
                        # its numbers in a dependency.  A generator that moved says so whatever the program is.)
                        if any(not torch.equal(a, b) for a, b in zip(before, after)):
                            raise ValueError(
                                f"'{name}' has a random initialisation program the planner cannot fold "
                                f"(first unfusable op: {info['first_unfusable_op'] or 'n/a'}); a plan cannot embed it")
                    else:
                        real = t
                    nbytes = real.numel() * real.element_size()
                    if nbytes > max_embedded_bytes:
                        raise ValueError(
                            f"'{name}' ({nbytes} bytes) has an initialisation program the planner cannot fold "
                            f"(first unfusable op: {info['first_unfusable_op'] or 'n/a'}) and is too large to embed")
                    raw = real.detach().cpu().contiguous().reshape(-1).view(torch.uint8).numpy().tobytes()
                    entries.append(PlanEntry(name, kind, list(real.shape), _DTYPE_NAMES[real.dtype],
                                             "value", value=base64.b64encode(raw).decode(),
                                             requires_grad=bool(real.requires_grad)))

        visit(module, "")
        return cls(entries)

    # --------------------------------------------------------------------------------------- io
    def save(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump({"format": FORMAT, "entries": [e.__dict__ for e in self.entries]}, f)

    @classmethod
    def load(cls, path: str) -> "InitPlan":
        with open(path) as f:
            doc = json.load(f)
        if doc.get("format") not in (FORMAT, "torchdistx_b200.InitPlan/1"):
            raise ValueError(f"{path}: not an {FORMAT} file")
        entries = []
        for d in doc["entries"]:
            d["epilogue"] = [tuple(e) for e in d.get("epilogue", [])]
            for g in d.get("segments", []):
                g["epilogue"] = [tuple(e) for e in g.get("epilogue", [])]
            entries.append(PlanEntry(**d))
        return cls(entries)

    @property
    def num_params(self) -> int:
        n = 0
        for e in self.entries:
            if e.kind == "param" and e.source != "alias":
                k = 1
                for s in e.sizes:
                    k *= s
                n += k
        return n

    # ------------------------------------------------------------------------------ descriptors
    def descriptors(self, seed: int, offset: int, shard: Optional[Tuple[int, int]] = None):
        """The plan as the C ABI sees it, without a device: ``([(entry, sizes, descs)], offset after)``
        for the tensors :meth:`materialize` would build from generator state ``(seed, offset)`` on
        this rank.  ``descs`` address each tensor from byte 0 (``dst`` = byte offset inside it); value
        and alias entries have none.  A host that is not Python rebuilds exactly this table from the
        JSON file; tests evaluate it with the stream's CPU restatement."""
        assigned: Dict[int, int] = {}
        table = []
        for e in self.entries:
            if e.source in ("alias", "value"):  # (a by-value parameter is chunked like any other: value_tensor)
                table.append((e, shard_range(e, shard)[2] if e.source == "value" else list(e.sizes), []))
                continue
            begin, count, sizes = shard_range(e, shard)
            pass_offset, offset = assign_pass_offsets(e, assigned, offset)
            table.append((e, sizes, entry_descriptors(e, 0, begin, count, seed, pass_offset)))
        return table, offset

    # ------------------------------------------------------------------------------ materialise
    def materialize(self, device="cuda", shard: Optional[Tuple[int, int]] = None,
                    generator: Optional[torch.Generator] = None,
                    into: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """Builds every tensor of the plan on `device` through libtdx_init and returns
        ``{name: tensor}`` (parameters as ``torch.nn.Parameter``).  Consumes the CUDA generator
        exactly like ``materialize_module`` would.

        `into`: write into caller-owned CUDA buffers instead of allocating (``{name: tensor}``; each
        must be contiguous, of the plan's dtype and hold at least this rank's elements).  This is
        the FSDP hand-off of SURVEY 8f.1: after ``fully_shard(meta_model)`` / ``to_empty`` the
        local shards FSDP owns are initialised in place, no unsharded tensor, no copy
        (:func:`init_sharded_module`).
        """
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("InitPlan.materialize drives the CUDA kernels; there is no CPU path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        gen = generator or torch.cuda.default_generators[device.index]
        seed, offset = gen.initial_seed(), gen.get_offset()
        out: Dict[str, torch.Tensor] = {}
        descs = []
        assigned: Dict[int, int] = {}  # RNG pass identity -> the offset it was given
        with torch.cuda.device(device):
            for e in self.entries:
                if e.source == "alias":
                    out[e.name] = out[e.alias_of]
                    continue
                dtype = _DTYPES[e.dtype]
                target = None if into is None else into.get(e.name)
                if e.source == "value":
                    t = value_tensor(e, shard).to(device)
                    if target is not None:
                        target.copy_(t)
                        t = target
                    out[e.name] = t if target is not None else _wrap(t, e)
                    continue
                begin, count, sizes = shard_range(e, shard)
                if target is not None:
                    if not (target.is_cuda and target.is_contiguous() and target.dtype == dtype
                            and target.numel() >= count):
                        raise ValueError(f"'{e.name}': `into` buffer must be a contiguous CUDA {dtype} tensor "
                                         f"with at least {count} elements")
                    t = target
                else:
                    t = torch.empty(sizes, dtype=dtype, device=device)
                pass_offset, offset = assign_pass_offsets(e, assigned, offset)
                descs += entry_descriptors(e, t.data_ptr(), begin, count, seed, pass_offset)
                out[e.name] = t if target is not None else _wrap(t, e)
            if descs:
                ws_bytes = _cabi.prepare(descs)  # exactly what this table needs (not the ~1.6 MB upper bound)
                if ws_bytes:
                    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
                    _cabi.submit(ws.data_ptr(), ws_bytes, torch.cuda.current_stream(device).cuda_stream)
            gen.set_offset(offset)
        return out

    def apply(self, module: Module, **kw) -> None:
        """Materialises the plan into `module` (e.g. one built under ``torch.device("meta")``):
        every parameter / buffer named in the plan is replaced."""
        tensors = self.materialize(**kw)
        for name, t in tensors.items():
            owner, _, key = name.rpartition(".")
            mod = module.get_submodule(owner) if owner else module
            group = mod._parameters if key in mod._parameters else mod._buffers
            group[key] = t


def _wrap(t: torch.Tensor, e: PlanEntry) -> torch.Tensor:
    if e.kind == "param":
        return torch.nn.Parameter(t, requires_grad=e.requires_grad)
    return t


def init_sharded_module(module: Module, plan: InitPlan, rank: int, world: int, device="cuda") -> None:
    """FSDP2 hand-off: `module` has been through ``fully_shard`` (its parameters are
    ``DTensor``s sharded on dim 0 over `world` ranks, e.g. built on the meta device and
    ``to_empty(device=...)``'d).  Every rank fills the local shards FSDP owns, in place, with its
    slice of the plan -- no full tensor, no chunk/clone.  All ranks must hold the same generator
    state (:func:`torchdistx_b200.parallel.sync_rng`)."""
    into: Dict[str, torch.Tensor] = {}
    for name, t in list(module.named_parameters()) + list(module.named_buffers()):
        local = t.to_local() if hasattr(t, "to_local") else t
        if hasattr(local, "data"):
            local = local.data
        into[name] = local
    with torch.no_grad():
        plan.materialize(device=device, shard=(rank, world), into=into)
