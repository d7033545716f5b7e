#!/usr/bin/env python
"""Multi-GPU sanity of plconv.Trainer under NCCL (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/trainer_ddp_check.py
Checks: parameters identical on every rank after training (same averaged gradients, same clip, same Adam), a NaN batch
on ONE rank is skipped by ALL ranks without a host sync, the loss goes down, sequences/s."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import plconv  # noqa: E402
from plconv.parallel import init_distributed  # noqa: E402


def main():
    rank, world, local = init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(100 + rank)                      # different init per rank: the Trainer must broadcast rank 0's
    cfg = plconv.TrainerConfig(hidden_dims=[16, 32], lu_channels=5, scale_factor=4, epochs=4, mode="bf16")
    tr = plconv.Trainer(cfg, device=dev)
    B = 4
    data = plconv.trainer.SyntheticRainBatches(6, B, 5, 15, 12, 4, 5, n_stations=30, seed=7 + rank)
    if rank == world - 1:                              # poison one batch on the last rank only
        bad = list(data.batches[2])
        bad[0] = bad[0].clone()
        bad[0][0, 0, 0, 0, 0] = float("nan")
        data.batches[2] = tuple(t.pin_memory() for t in bad)
    t0 = time.perf_counter()
    hist = tr.fit(data)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    flat = torch.cat([p.detach().flatten() for p in tr.model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([float(torch.equal(flat, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        steps = cfg.epochs * len(data)
        print(f"world {world}: params identical on all ranks: {bool(same.item())}; skipped (rank-steps) {tr.skipped} "
              f"(expected {cfg.epochs * world}); loss {hist['total_loss'][0]:.3f} -> {hist['total_loss'][-1]:.3f}; "
              f"{world * B * steps / dt:.0f} sequences/s incl. first-step warm-up")
        assert bool(same.item()) and tr.skipped == cfg.epochs * world and hist["total_loss"][-1] < hist["total_loss"][0]
        assert all(torch.isfinite(p).all() for p in tr.model.parameters())
        print("OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
