"""One optimizer step of the recurrence under data parallelism, mirroring the reference loop
(src/training/trainer.py:286-315): zero_grad -> forward -> loss -> NaN-skip -> backward -> clip(0.5) -> Adam step,
with the gradient all-reduce inserted between backward and clipping.

On CUDA the step never synchronises with the host: the NaN-skip is a device flag (max-reduced over the ranks) that
fused Adam consumes as ``found_inf`` -- the same device-side skip `Trainer.train_step` uses -- so the host runs ahead
of the GPU and kernel launches of step i+1 are queued while step i still computes."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch

from . import _lib
from .parallel import GradReducer, all_ranks_finite, nonfinite_flag


class TrainStep:
    def __init__(self, model: torch.nn.Module, groups, lr: float = 5e-4, grad_clip_norm: float = 0.5,
                 process_group=None, betas=(0.9, 0.999)):
        self.model = model
        self.reducer = GradReducer(groups, process_group)
        self.reducer.attach_cell_sinks(model)                    # per-layer hand-over inside the fused rollouts
        self.params = [p for b in self.reducer.buckets for p in b["params"]]
        self.on_cuda = all(p.is_cuda for p in self.params)
        # trainer.py:153-158 (Adam, lr from config); fused on CUDA so that it honours the device-side skip flag
        # capturable: the step counter lives on the device, so the whole step can be recorded in a CUDA graph (GraphedStep)
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=betas, fused=self.on_cuda, capturable=self.on_cuda)
        self.clip = grad_clip_norm                               # trainer.py:311-314 (0.5)
        self.pg = process_group
        self.skipped = 0                                         # host-visible count (CPU path only)
        self.skipped_dev = None                                  # device-side count (CUDA path), read it when convenient

    def zero_grad(self):
        self.reducer.zero_grad()                                 # trainer.py:290

    def backward_and_step(self, loss: torch.Tensor) -> Optional[torch.Tensor]:
        """NaN-skip -> backward -> all-reduce -> clip -> Adam for an already computed loss."""
        if not self.on_cuda:
            if not all_ranks_finite(loss, self.pg):              # trainer.py:306-308, made rank-consistent
                self.skipped += 1
                return None
            loss.backward()                                      # trainer.py:310
            self.reducer.finish()
            torch.nn.utils.clip_grad_norm_(self.params, self.clip)
            self.opt.step()
            return loss.detach()
        bad = nonfinite_flag(loss, self.pg)                      # device flag; every rank skips together
        loss.backward()                                          # trainer.py:310 (BPTT -> plc_cell_bwd per step)
        self.reducer.finish()                                    # mean of gradients over ranks
        torch.nn.utils.clip_grad_norm_(self.params, self.clip)   # trainer.py:311-314, AFTER the all-reduce
        self.opt.grad_scale, self.opt.found_inf = None, bad.reshape(())
        self.opt.step()                                          # trainer.py:315; skipped on the device if bad
        del self.opt.grad_scale, self.opt.found_inf
        if self.skipped_dev is None:
            self.skipped_dev = torch.zeros_like(bad)
        self.skipped_dev.add_(bad)                               # in place: keeps counting across CUDA-graph replays
        return loss.detach()

    def __call__(self, forward_loss: Callable[[], torch.Tensor]) -> Optional[torch.Tensor]:
        self.zero_grad()
        return self.backward_and_step(forward_loss())            # trainer.py:297-304


class GraphedStep:
    """A whole training step -- forward, loss, BPTT, gradient all-reduce, clip, Adam -- recorded ONCE as a CUDA graph and
    replayed per batch.

    Why: at the 8-GPU shard of cfg3 (8 sequences per GPU) a cell kernel runs ~55 us and a step is ~240 library launches
    plus ~150 small torch kernels (optimizer, clip, loss); issued from Python the GPU idles between launches (measured:
    12.6 % of the step at batch 8 against 4.3 % at batch 64).  A graph replay has no per-launch host work at all, so strong scaling no longer
    depends on the host keeping ahead.  Everything the step does is already capture-safe: no host synchronisation
    (device-side NaN-skip handed to fused Adam as ``found_inf``), capturable Adam, tensor maps passed as kernel
    parameters, state rings allocated from the graph's private pool (same addresses on every replay), NCCL all-reduces
    recorded on their own stream with event edges.

    ``step_fn(*inputs) -> tensor`` must be shape-static.  ``warmup`` eager calls run first on a side stream (lazy
    optimizer state, NCCL communicators, smem attributes); they are REAL optimizer steps on ``example_inputs``.
    ``__call__`` copies the batch into the static input buffers, replays, and returns the static output tensor (valid
    until the next call)."""

    def __init__(self, step_fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        if not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedStep needs CUDA tensors (CUDA graphs)")
        self.step_fn = step_fn
        self.static_in = [t.detach().clone() for t in example_inputs]
        dev = self.static_in[0].device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                step_fn(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        # the capture allocates the step's buffers (state rings: tens of GB at radar scale) from the graph's PRIVATE pool;
        # blocks the eager warm-up left in the caching allocator cannot be reused there, so hand them back first
        torch.cuda.empty_cache()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = step_fn(*self.static_in)
        self.eager_steps = max(warmup, 1)            # optimizer steps taken so far (capturing records, it does not execute)
        _lib.invalidate_packed_weights()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        # the replay repacked and then updated the weights behind the Python-side caches' backs
        _lib.invalidate_packed_weights()
        return self.static_out

    def close(self) -> None:
        """Drop the recorded graph (and with it the captured NCCL launches).  Call before
        ``torch.distributed.destroy_process_group()``: tearing a communicator down while a live graph still holds its
        kernels can block in NCCL's destroy path."""
        if self.graph is not None:
            dev = self.static_in[0].device
            torch.cuda.synchronize(dev)
            self.graph.reset()
            self.graph = None
            self.static_out = None
            torch.cuda.empty_cache()                 # return the private pool to the driver
