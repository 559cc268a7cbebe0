"""Random initialisation programs (TEST INFRASTRUCTURE: tests/test_planner_fuzz_cpu.py and the
reference-side driver oracle/ref_fuzz_driver.py).  Pure PyTorch: imports neither torchdistx_b200 nor
the compiled reference, so both sides of a differential run build exactly the same programs."""
import math

import torch
from torch import nn

CONSTS = [0.5, 2.0, -1.5, 0.02, 3.0, 0.25, 1.0, -0.125]
STEPS = ["mul", "add", "clamp", "reinit_u", "reinit_n", "row_zero", "slice_normal", "slice_fill", "slice_mul",
         "oop_affine", "clone", "detach", "zero", "div", "sub", "neg", "cast16", "castback", "noop_cast"]


def gen_program(r):
    rows, cols = r.randint(32, 96), r.choice([8, 16, 24, 32])
    ctor = r.choice(["empty", "empty", "empty", "zeros", "ones", "full", "randn", "rand"])
    steps = [("ctor", ctor, rows, cols, r.choice(CONSTS))]
    if ctor == "empty":
        steps.append(("init", r.choice(["uniform", "normal", "fill", "trunc", "kaiming", "xavier"]),
                      r.choice([(-0.1, 0.1), (0.0, 1.0), (-0.05, 0.03)]), r.choice([(0.0, 0.02), (1.0, 0.5), (0.0, 1.0)]),
                      r.choice(CONSTS)))
    for _ in range(r.randint(0, 4)):
        a, b = sorted([r.randint(0, rows), r.randint(0, rows)])
        if a == b:
            b = min(rows, a + 1)
            a = b - 1
        steps.append((r.choice(STEPS), a, b, r.choice(CONSTS), r.choice(CONSTS), r.choice([torch.bfloat16, torch.float16])))
    return steps


VIEW_STEPS = ["col_zero", "flat_fill", "flat_normal", "t_row_fill", "copy_from_rng", "copy_from_const", "copy_slice",
              "view_mul", "unsq_add", "reshape_fill", "narrow_uniform", "select_col_normal", "chunk_fill", "expand_copy",
              "fill_tensor", "add_alpha", "mul_tensor0d"]


def apply_step(t, st):
    k = st[0]
    if k == "ctor":
        _, c, rows, cols, v = st
        return {"empty": lambda: torch.empty(rows, cols), "zeros": lambda: torch.zeros(rows, cols),
                "ones": lambda: torch.ones(rows, cols), "full": lambda: torch.full((rows, cols), v),
                "randn": lambda: torch.randn(rows, cols), "rand": lambda: torch.rand(rows, cols)}[c]()
    if k == "init":
        _, kind, (lo, hi), (m, s), v = st
        if kind == "uniform":
            t.uniform_(lo, hi)
        elif kind == "normal":
            t.normal_(m, s)
        elif kind == "fill":
            t.fill_(v)
        elif kind == "trunc":
            nn.init.trunc_normal_(t, mean=m, std=s, a=m - 2 * s, b=m + 1.5 * s)
        elif kind == "kaiming":
            nn.init.kaiming_uniform_(t, a=math.sqrt(5))
        else:
            nn.init.xavier_normal_(t)
        return t
    _, a, b, c, d, x = st  # x: a 16-bit dtype (base vocabulary) or a column index (view vocabulary)
    rows, cols = t.shape
    if k == "mul":
        t.mul_(c)
    elif k == "add":
        t.add_(c)
    elif k == "div":
        t.div_(c)
    elif k == "sub":
        t.sub_(c)
    elif k == "neg":
        t.neg_()
    elif k == "clamp":
        t.clamp_(min(c, d), max(c, d))
    elif k == "reinit_u":
        t.uniform_(-abs(c), abs(c))
    elif k == "reinit_n":
        t.normal_(0.0, abs(c))
    elif k == "row_zero":
        t[a].zero_()
    elif k == "slice_normal":
        t[a:b].normal_(d, abs(c))
    elif k == "slice_fill":
        t[a:b].fill_(c)
    elif k == "slice_mul":
        t[a:b].mul_(c)
    elif k == "oop_affine":
        t = t * c + d
    elif k == "clone":
        t = t.clone()
    elif k == "detach":
        t = t.detach()
    elif k == "zero":
        t.zero_()
    elif k == "cast16":
        t = t.to(x)
    elif k == "castback":
        t = t.to(torch.float32)
    elif k == "noop_cast":
        t = t.to(t.dtype)
    # writes through views of every kind (contiguous ranges fold, strided ones must fall back), copies
    elif k == "col_zero":
        t[:, x].zero_()
    elif k == "flat_fill":
        t.view(-1)[a * cols + 3: b * cols].fill_(c)
    elif k == "flat_normal":
        t.view(-1)[a * cols: b * cols].normal_(d, abs(c))
    elif k == "t_row_fill":
        t.t()[x].fill_(c)
    elif k == "copy_from_rng":
        t.copy_(torch.empty(rows, cols, dtype=t.dtype).uniform_(-abs(c), abs(c)))
    elif k == "copy_from_const":
        t.copy_(torch.full((rows, cols), c, dtype=t.dtype))
    elif k == "copy_slice":
        t[a:b].copy_(torch.full((b - a, cols), d, dtype=t.dtype))
    elif k == "view_mul":
        t.view(rows * cols).mul_(c)
    elif k == "unsq_add":
        t.unsqueeze(0).add_(c)
    elif k == "reshape_fill":
        t.reshape(-1, cols)[a:b].fill_(c)
    elif k == "narrow_uniform":
        t.narrow(0, a, b - a).uniform_(-abs(c), abs(c))
    elif k == "select_col_normal":
        t.select(1, x).normal_(0, abs(c))
    elif k == "chunk_fill":
        t.chunk(2, 0)[1].fill_(c)
    elif k == "expand_copy":
        t.copy_(torch.full((1, cols), c, dtype=t.dtype).expand(rows, cols))
    elif k == "fill_tensor":
        t.fill_(torch.tensor(c))
    elif k == "add_alpha":
        t.add_(c, alpha=d)
    elif k == "mul_tensor0d":
        t.mul_(torch.tensor(c))
    else:
        raise KeyError(k)
    return t


def run_program(steps):
    t = None
    for st in steps:
        t = apply_step(t, st)
    return t


class Holder(nn.Module):
    def __init__(self, progs):
        super().__init__()
        for i, p in enumerate(progs):
            setattr(self, f"t{i}", nn.Parameter(run_program(p)))


def gen_view_program(r):
    steps = gen_program(r)
    rows, cols = steps[0][2], steps[0][3]
    for _ in range(r.randint(1, 3)):
        a, b = sorted([r.randint(0, rows), r.randint(0, rows)])
        if a == b:
            b = min(rows, a + 1)
            a = b - 1
        pos = r.randint(2 if steps[0][1] == "empty" else 1, len(steps))  # (never before the tensor has values)
        steps.insert(pos, (r.choice(VIEW_STEPS), a, b, r.choice(CONSTS), r.choice(CONSTS), r.randint(0, cols - 1)))
    return steps


ROWS, COLS = 48, 16
LINKS = ["clone", "affine", "copy_into_empty", "copy_into_rng", "slice_copy", "detach_clone_mul", "cast16", "mul_by_const_tensor"]


def gen_linked(r):
    """Two or three tensors of one shape; a later one may START as a function of an earlier one's
    state part-way through ITS program (a reader of an intermediate state), and both go on changing."""
    n = r.randint(2, 3)
    progs = []
    for _ in range(n):
        base = gen_program(r)
        base[0] = ("ctor", base[0][1], ROWS, COLS, base[0][4])
        fixed = []
        for st in base:
            if st[0] not in ("ctor", "init"):
                b = min(max(st[2], 1), ROWS)
                a = min(st[1], b - 1)
                st = (st[0], a, b) + st[3:]
            fixed.append(st)
        progs.append(fixed)
    links = {}
    for j in range(1, n):
        if r.random() < 0.8:
            i = r.randrange(0, j)
            first = 2 if progs[i][0][1] == "empty" else 1
            links[j] = (i, r.randint(first, len(progs[i])), r.choice(LINKS), r.choice(CONSTS), r.choice(CONSTS), r.randint(1, ROWS - 1))
    return progs, links


def run_linked(progs, links):
    wanted = {}
    for j, (i, cut, kind, c, d, k) in links.items():
        wanted.setdefault((i, cut), []).append((j, kind, c, d, k))
    start, out = {}, {}
    for idx, steps in enumerate(progs):
        t = start.get(idx)
        if t is not None:
            steps = [st for st in steps if st[0] not in ("ctor", "init")]
        for n_done, st in enumerate(steps, start=1):
            t = apply_step(t, st)
            for j, kind, c, d, k in (wanted.get((idx, n_done), []) if idx not in start else []):
                if kind == "clone":
                    start[j] = t.clone()
                elif kind == "affine":
                    start[j] = t * c + d
                elif kind == "copy_into_empty":
                    start[j] = torch.empty(ROWS, COLS, dtype=t.dtype).copy_(t)
                elif kind == "copy_into_rng":
                    start[j] = torch.randn(ROWS, COLS).to(t.dtype).copy_(t)
                elif kind == "slice_copy":
                    x = torch.zeros(ROWS, COLS, dtype=t.dtype)
                    x[:k].copy_(t[:k])
                    start[j] = x
                elif kind == "detach_clone_mul":
                    start[j] = t.detach().clone().mul_(c)
                elif kind == "cast16":
                    start[j] = t.to(torch.bfloat16)
                else:
                    start[j] = t * torch.full((ROWS, COLS), c, dtype=t.dtype)
        out[idx] = t
    return out


class LinkedHolder(nn.Module):
    def __init__(self, progs, links):
        super().__init__()
        for i, t in sorted(run_linked(progs, links).items()):
            setattr(self, f"t{i}", nn.Parameter(t))




# ---- the Parameter / `.data` / nn.init / Module.to idioms of model code (HF `_init_weights`) ----------
DATA_STEPS = ["data_normal", "data_uniform", "data_zero_row", "data_fill", "data_assign_affine", "data_assign_clone", "data_copy_const",
              "nograd_mul", "nograd_row_fill", "data_mul", "data_slice_normal", "requires_grad_false", "data_to16", "data_float",
              "module_to16", "module_float", "module_to_cpu", "data_clamp", "init_trunc", "init_kaiming", "init_zeros", "init_constant", "init_normal", "init_xavier_u"]

class Cell(nn.Module):
    def __init__(self, prog, dsteps):
        super().__init__()
        self.p = nn.Parameter(run_program(prog))
        rows, cols = self.p.shape
        for (k, a, b, c, d) in dsteps:
            a = min(a, rows - 1); b = max(a + 1, min(b, rows))
            p = self.p
            if k == "data_normal": p.data.normal_(d, abs(c))
            elif k == "data_uniform": p.data.uniform_(-abs(c), abs(c))
            elif k == "data_zero_row": p.data[a].zero_()
            elif k == "data_fill": p.data.fill_(c)
            elif k == "data_assign_affine": p.data = p.data * c + d
            elif k == "data_assign_clone": p.data = p.data.clone()
            elif k == "data_copy_const": p.data.copy_(torch.full(tuple(p.shape), c, dtype=p.dtype))
            elif k == "nograd_mul":
                with torch.no_grad(): p.mul_(c)
            elif k == "nograd_row_fill":
                with torch.no_grad(): p[a:b].fill_(d)
            elif k == "data_mul": p.data.mul_(c)
            elif k == "data_slice_normal": p.data[a:b].normal_(0.0, abs(c))
            elif k == "requires_grad_false": p.requires_grad_(False)
            elif k == "data_to16": p.data = p.data.to(torch.bfloat16)
            elif k == "data_float": p.data = p.data.float()
            elif k == "module_to16": self.to(torch.bfloat16)
            elif k == "module_float": self.float()
            elif k == "module_to_cpu": self.to("cpu")
            elif k == "data_clamp": p.data.clamp_(min(c, d), max(c, d))
            elif k == "init_trunc": nn.init.trunc_normal_(p, std=abs(c), a=-2 * abs(c), b=2 * abs(c))
            elif k == "init_kaiming": nn.init.kaiming_uniform_(p, a=5 ** 0.5)
            elif k == "init_zeros": nn.init.zeros_(p)
            elif k == "init_constant": nn.init.constant_(p, c)
            elif k == "init_normal": nn.init.normal_(p, d, abs(c))
            elif k == "init_xavier_u": nn.init.xavier_uniform_(p)

def gen_cell(r):
    prog = gen_program(r)
    rows = prog[0][2]
    ds = []
    for _ in range(r.randint(1, 5)):
        a, b = sorted([r.randint(0, rows), r.randint(0, rows)])
        ds.append((r.choice(DATA_STEPS), a, b, r.choice(CONSTS), r.choice(CONSTS)))
    return prog, ds


def differential_script(seed):
    """The script both sides of a differential run (this engine / the compiled reference) build for
    `seed`: linked tensors, view and copy steps sprinkled in, random draws included."""
    import random

    r = random.Random(40_000 + seed)
    progs, links = gen_linked(r)
    for p in progs:
        for _ in range(r.randint(0, 2)):
            b = r.randint(1, ROWS)
            first = 2 if p[0][1] == "empty" else 1
            p.insert(r.randint(first, len(p)), (r.choice(VIEW_STEPS), r.randint(0, b - 1), b, r.choice(CONSTS), r.choice(CONSTS),
                                                r.randint(0, COLS - 1)))
    return progs, links
