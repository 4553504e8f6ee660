"""Adam over flat fp32 arenas (one kernel per optimiser step) + data-parallel gradient
all-reduce of the same arena over NCCL.

Replaces torch.optim.Adam as used by reference train.py:94-116 (defaults: eps 1e-8, no weight
decay, no amsgrad).  Parameters, gradients and both moments of one network live in four
contiguous buffers; `param.data` / `param.grad` are views into them, so autograd accumulates
straight into the all-reduce bucket and the update is a single 128-bit-vectorised pass."""

from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import kernels as K
from . import ops

_ALIGN = 64  # elements; keeps every tensor 256-byte aligned inside the arena


# OTM_DDP_OVERLAP=0: every gradient all-reduce is waited for where it is issued, i.e. the NCCL
# kernels never run beside compute kernels.  Overlap hides the transfer but NCCL's CTAs take SMs
# from the persistent one-CTA-per-SM kernels running beside them (each of those then needs a
# second wave); which side wins depends on the world size.
OVERLAP = os.environ.get("OTM_DDP_OVERLAP", "1") != "0"


class GradArena:
    """Flat parameter / gradient arenas of one network plus the data-parallel all-reduce.
    Pure tensor plumbing (device agnostic): `param.data` and `param.grad` become views of two
    contiguous fp32 buffers, so autograd accumulates straight into the all-reduce bucket."""

    def __init__(self, params, *, process_group=None, data_parallel: bool | None = None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("empty parameter list")
        dev = self.params[0].device
        self.offsets = []
        total = 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = total
        self.param_arena = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad_arena = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.param_arena[off : off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self._attach_grads()
        self.group = process_group
        if data_parallel is None:
            data_parallel = (dist.is_available() and dist.is_initialized()
                             and dist.get_world_size(process_group) > 1)
        self.data_parallel = data_parallel
        self.world = dist.get_world_size(process_group) if data_parallel else 1
        self._pending: list = []                   # async NCCL works since the last wait
        self._issued: list[tuple[int, int]] = []   # their (disjoint) element ranges
        for p in self.params:
            p._otm_owner = id(self)  # tags this optimiser's entries of the weight-pack cache

    def _attach_grads(self):
        for p, off in zip(self.params, self.offsets):
            p.grad = self.grad_arena[off : off + p.numel()].view_as(p)

    def zero_grad(self, set_to_none: bool = False):
        """One memset over the arena; gradients stay views of it (set_to_none is accepted for
        signature compatibility and ignored)."""
        self.grad_arena.zero_()
        base = self.grad_arena.data_ptr()
        if any(p.grad is None or p.grad.data_ptr() != base + 4 * off
               for p, off in zip(self.params, self.offsets)):
            self._attach_grads()

    def offset_of(self, param) -> int:
        """Arena offset (elements) of a parameter: bucket boundaries for the overlapped reduce."""
        for p, off in zip(self.params, self.offsets):
            if p is param:
                return off
        raise KeyError("parameter is not in this arena")

    def all_reduce_async(self, start: int = 0, end: int | None = None):
        """Sum-all-reduce elements [start, end) of the gradient arena (NCCL over NVLink on the
        GPU box) without blocking: a bucket can be launched as soon as its gradients are final,
        while the rest of backward still runs.  The 1/world averaging is folded into the Adam
        kernel's grad_scale.  Buckets issued between two waits must not overlap (an in-place sum
        applied twice would count a gradient twice); re-issuing a covered range is a no-op."""
        end = self.numel if end is None else end
        if not self.data_parallel or start >= end:
            return
        for lo, hi in self._issued:
            if lo <= start and end <= hi:
                return
            if start < hi and lo < end:
                raise ValueError(f"all-reduce bucket [{start},{end}) overlaps [{lo},{hi})")
        self._issued.append((start, end))
        work = dist.all_reduce(self.grad_arena[start:end], op=dist.ReduceOp.SUM, group=self.group,
                               async_op=True)
        if OVERLAP:
            self._pending.append(work)
        else:
            work.wait()  # the compute stream waits here: the collective runs alone on the GPU

    def _missing(self):
        gaps, pos = [], 0
        for lo, hi in sorted(self._issued):
            if lo > pos:
                gaps.append((pos, lo))
            pos = max(pos, hi)
        if pos < self.numel:
            gaps.append((pos, self.numel))
        return gaps

    def wait_all_reduce(self):
        """Reduce whatever part of the arena has not been issued yet, then wait for everything."""
        if not self.data_parallel:
            return
        for lo, hi in self._missing():
            self.all_reduce_async(lo, hi)
        for w in self._pending:
            w.wait()
        self._pending = []
        self._issued = []


def all_reduce_buckets(buckets):
    """Sum-all-reduce several arena ranges -- [(arena, start, end | None), ...], possibly of
    different optimisers -- as ONE coalesced collective (one NCCL group launch instead of one per
    arena: at 8 ranks a gradient all-reduce of a few MB costs ~0.25 ms of mostly fixed latency, so
    the count matters more than the bytes).  Same bookkeeping as GradArena.all_reduce_async: every
    involved arena waits for the shared work in wait_all_reduce()."""
    todo = []
    for arena, start, end in buckets:
        end = arena.numel if end is None else end
        if not arena.data_parallel or start >= end:
            continue
        if any(lo <= start and end <= hi for lo, hi in arena._issued):
            continue
        for lo, hi in arena._issued:
            if start < hi and lo < end:
                raise ValueError(f"all-reduce bucket [{start},{end}) overlaps [{lo},{hi})")
        todo.append((arena, start, end))
    if not todo:
        return
    group = todo[0][0].group
    if len(todo) == 1:
        a, lo, hi = todo[0]
        a.all_reduce_async(lo, hi)
        return
    from torch.distributed.distributed_c10d import _coalescing_manager

    with _coalescing_manager(group=group, async_ops=True) as cm:
        for a, lo, hi in todo:
            dist.all_reduce(a.grad_arena[lo:hi], op=dist.ReduceOp.SUM, group=group)
    for a, lo, hi in todo:
        a._issued.append((lo, hi))
    if OVERLAP:
        for a in {id(t[0]): t[0] for t in todo}.values():
            a._pending.append(cm)
    else:
        cm.wait()


class FlatAdam(GradArena):
    def __init__(self, params, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, *,
                 process_group=None, data_parallel: bool | None = None):
        super().__init__(params, process_group=process_group, data_parallel=data_parallel)
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam needs CUDA parameters (no CPU fallback)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.steps = 0
        ops.invalidate_packs()

    def step(self, reduced: bool = False):
        """`reduced=True`: the caller has already completed the gradient all-reduce (the graph
        engine keeps NCCL outside its captured segments)."""
        if not reduced:
            self.wait_all_reduce()
        self.steps += 1
        self.step_dev += 1
        K.adam(self.param_arena, self.grad_arena, self.exp_avg, self.exp_avg_sq, self.step_dev,
               self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.world)
        ops.invalidate_packs(owner=id(self))  # only THIS network's staged weights are stale

    # torch.optim.Adam-compatible checkpoint layout (reference evaluation.py:248-263)
    def state_dict(self):
        state = {}
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            n = p.numel()
            state[i] = {
                "step": torch.tensor(float(int(self.step_dev.item()))),
                "exp_avg": self.exp_avg[off : off + n].view_as(p).clone(),
                "exp_avg_sq": self.exp_avg_sq[off : off + n].view_as(p).clone(),
            }
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        with torch.no_grad():
            for i, (p, off) in enumerate(zip(self.params, self.offsets)):
                st = sd["state"].get(i)
                if st is None:
                    continue
                n = p.numel()
                self.exp_avg[off : off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off : off + n].copy_(st["exp_avg_sq"].reshape(-1))
                self.steps = int(st["step"])
            self.step_dev.fill_(self.steps)
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])


def shutdown_process_group(timeout_s: float = 20.0) -> None:
    """barrier + destroy_process_group that cannot hang the process: communicators that were
    captured into CUDA graphs have been seen to block in ncclCommDestroy even after the graphs
    were released, so the teardown runs in a helper thread and, if it does not return in time,
    the process exits immediately with status 0 (all results have been written by then)."""
    import os
    import sys
    import threading

    if not (dist.is_available() and dist.is_initialized()):
        return
    dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout_s)
    if t.is_alive():
        os._exit(0)
