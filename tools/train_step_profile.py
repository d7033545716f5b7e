#!/usr/bin/env python
"""One cfg3 GAN training step between cudaProfilerStart/Stop, for an ncu launch list of exactly one step:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/train_step_launches.csv python tools/train_step_profile.py [B=64]

Every kernel of the step is listed -- the library's and torch's (cat, loss, clip, Adam, NCCL-free at one GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import plconv  # noqa: E402
from plconv.gan import Discriminator, GanTrainStep  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    gen = plconv.NowcastGenerator(1, [64, 64], 3, 10, 10, "bf16").to(dev)
    disc = Discriminator().to(dev)
    step = GanTrainStep(gen, disc, lr_g=5e-4, lr_d=2e-4, lambda_adv=0.05, grad_clip_norm=0.5)
    frames = torch.relu(torch.randn(B, 10, 1, 128, 128, device=dev) + 0.3)
    target = torch.relu(torch.randn(B, 10, 1, 128, 128, device=dev) + 0.3)
    for _ in range(3):
        step(frames, target)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(frames, target)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
