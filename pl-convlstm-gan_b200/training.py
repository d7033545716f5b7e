"""One optimizer step of the recurrence under data parallelism, mirroring the reference loop
(src/training/trainer.py:286-315): zero_grad -> forward -> loss -> NaN-skip -> backward -> clip(0.5) -> Adam step,
with the gradient all-reduce inserted between backward and clipping."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .parallel import GradReducer, all_ranks_finite


class TrainStep:
    def __init__(self, model: torch.nn.Module, groups, lr: float = 5e-4, grad_clip_norm: float = 0.5,
                 process_group=None):
        self.model = model
        self.reducer = GradReducer(groups, process_group)
        self.params = [p for b in self.reducer.buckets for p in b["params"]]
        self.opt = torch.optim.Adam(self.params, lr=lr)          # trainer.py:153-158 (Adam, lr from config)
        self.clip = grad_clip_norm                               # trainer.py:311-314 (0.5)
        self.pg = process_group
        self.skipped = 0

    def __call__(self, forward_loss: Callable[[], torch.Tensor]) -> Optional[torch.Tensor]:
        self.reducer.zero_grad()                                 # trainer.py:290
        loss = forward_loss()                                    # trainer.py:297-304
        if not all_ranks_finite(loss, self.pg):                  # trainer.py:306-308, made rank-consistent
            self.skipped += 1
            return None
        loss.backward()                                          # trainer.py:310 (BPTT -> plc_cell_bwd per step)
        self.reducer.finish()                                    # mean of gradients over ranks
        torch.nn.utils.clip_grad_norm_(self.params, self.clip)   # trainer.py:311-314, AFTER the all-reduce
        self.opt.step()                                          # trainer.py:315
        return loss.detach()
