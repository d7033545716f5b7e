"""One optimizer step of the recurrence under data parallelism, mirroring the reference loop
(src/training/trainer.py:286-315): zero_grad -> forward -> loss -> NaN-skip -> backward -> clip(0.5) -> Adam step,
with the gradient all-reduce inserted between backward and clipping.

On CUDA the step never synchronises with the host: the NaN-skip is a device flag (max-reduced over the ranks) that
fused Adam consumes as ``found_inf`` -- the same device-side skip `Trainer.train_step` uses -- so the host runs ahead
of the GPU and kernel launches of step i+1 are queued while step i still computes."""
from __future__ import annotations

from typing import Callable, Optional

import torch

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
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=betas, fused=self.on_cuda)
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
        self.skipped_dev = bad if self.skipped_dev is None else self.skipped_dev + bad
        return loss.detach()

    def __call__(self, forward_loss: Callable[[], torch.Tensor]) -> Optional[torch.Tensor]:
        self.zero_grad()
        return self.backward_and_step(forward_loss())            # trainer.py:297-304
