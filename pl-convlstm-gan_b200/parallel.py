"""Data-parallel plumbing: one process per GPU, batch sharding, bucketed gradient all-reduce overlapped with BPTT.

The reference has no distributed code at all ("Distributed comm backend: absent").  The path
shards naturally by batch (no cross-sample op, no BatchNorm: generator.py:156-171), so the only exchange is ONE
all-reduce of the (tiny: <= 14 MB) gradients per optimizer step.  `GradReducer` launches it per bucket as soon as
the bucket's last gradient has been accumulated -- for a cell that is when its t = 0 BPTT step retires -- so the
collective overlaps the BPTT of the layers below.  NCCL over NVLink on GPU; gloo on CPU for the tests.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> tuple:
    """Read RANK / WORLD_SIZE / LOCAL_RANK (torchrun contract); returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_batch(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous samples [lo, hi) of a global batch owned by `rank`."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


class GradReducer:
    """Bucketed gradient averaging across ranks, overlapped with the backward pass.

    buckets: list of parameter lists (default: one bucket per top-level group passed in `groups`, e.g. one per
    ConvLSTM cell).  Each bucket owns ONE flat fp32 buffer; every parameter's `.grad` is a view into it, so the
    collective needs no gather/scatter copies.  A post-accumulate-grad hook per parameter counts arrivals; when a
    bucket is complete its all-reduce is launched asynchronously (NCCL runs it on its own stream, ordered after the
    kernels already queued on the current stream).  `finish()` waits for all buckets and applies 1/world.
    """

    def __init__(self, groups: Sequence[Iterable[torch.nn.Parameter]], process_group=None):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets: List[dict] = []
        self._hooks = []
        self._sunk: List[torch.nn.Module] = []
        for grp in groups:
            params = [p for p in grp if p.requires_grad]
            if not params:
                continue
            n = sum(p.numel() for p in params)
            flat = torch.zeros(n, dtype=torch.float32, device=params[0].device)
            off = 0
            for p in params:
                if p.dtype != torch.float32:
                    raise ValueError("GradReducer expects fp32 master parameters")
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            b = {"params": params, "flat": flat, "pending": len(params), "handle": None}
            self.buckets.append(b)
            for p in params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))

    def attach_cell_sinks(self, module: torch.nn.Module) -> int:
        """Per-LAYER hand-over inside a fused rollout: `_StackRolloutFn` is one autograd node per stack, so the
        post-accumulate hooks above would only fire when the whole stack's BPTT has been queued.  With a sink attached,
        the rollout's backward gives a cell's dW / db to its bucket the moment that layer's t = 0 step is queued, and
        the bucket's all-reduce overlaps the BPTT steps still to run.  Returns the number of cells attached."""
        by_param = {id(p): b for b in self.buckets for p in b["params"]}
        n = 0
        for cell in module.modules():
            conv = getattr(cell, "conv", None)
            if not hasattr(cell, "_packed") or conv is None:
                continue
            ps = [p for p in (conv.weight, conv.bias) if p is not None]
            bs = {id(by_param.get(id(p))) for p in ps}
            if len(bs) != 1 or by_param.get(id(ps[0])) is None:
                continue                                     # not (entirely) in one of this reducer's buckets
            cell._grad_sink = self._make_sink(by_param[id(ps[0])])
            self._sunk.append(cell)
            n += 1
        return n

    def _make_sink(self, bucket):
        def sink(cell, gw, gb) -> bool:
            conv = cell.conv
            if conv.weight.grad is None or (gb is not None and conv.bias.grad is None):
                return False                                 # .grad detached from the bucket: let autograd accumulate
            conv.weight.grad.add_(gw)
            k = 1
            if gb is not None:
                conv.bias.grad.add_(gb)
                k = 2
            bucket["pending"] -= k
            if bucket["pending"] == 0:
                self._launch(bucket)
            return True
        return sink

    def _make_hook(self, bucket):
        def hook(_param):
            bucket["pending"] -= 1
            if bucket["pending"] == 0:
                self._launch(bucket)
        return hook

    def _launch(self, bucket):
        if self.world > 1:
            bucket["handle"] = dist.all_reduce(bucket["flat"], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def zero_grad(self):
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            b["handle"] = None
            for p in b["params"]:      # keep .grad pointing into the flat buffer (optimizers may have replaced it)
                if p.grad is None or p.grad.data_ptr() < b["flat"].data_ptr() or \
                        p.grad.data_ptr() >= b["flat"].data_ptr() + b["flat"].numel() * 4:
                    raise RuntimeError("parameter .grad was detached from its bucket; use reducer.zero_grad() only")

    def finish(self):
        """Wait for every bucket's collective and turn sums into means (call after backward, before clipping)."""
        for b in self.buckets:
            if b["pending"] != 0 and self.world > 1:
                # parameter unused in this step: its gradient contribution is zero, still reduce for consistency
                self._launch(b)
            if b["handle"] is not None:
                b["handle"].wait()
            if self.world > 1:
                b["flat"].div_(self.world)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for cell in self._sunk:
            cell._grad_sink = None
        self._sunk = []


def nonfinite_flag(value: torch.Tensor, process_group=None) -> torch.Tensor:
    """Device-side form of the NaN-skip (trainer.py:306-308): a float32 scalar tensor, 1 where ANY rank's `value` is
    not finite, else 0 -- max-reduced over the ranks, never read by the host.  Hand it to a fused optimizer as
    ``found_inf`` (the update is skipped on the device, as torch.amp.GradScaler does) so a step has no host sync."""
    bad = (~torch.isfinite(value.detach())).any().to(torch.float32).reshape(1)
    if dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=process_group)
    return bad


def all_ranks_finite(value: torch.Tensor, process_group=None) -> bool:
    """Collective-safe version of the reference's NaN-skip (`if torch.isnan(loss): continue`, trainer.py:306-308):
    every rank must take the same branch or the next collective deadlocks."""
    flag = torch.isfinite(value.detach()).all().to(torch.float32).reshape(1)
    if dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=process_group)
    return bool(flag.item() > 0.5)
