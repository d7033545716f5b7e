#!/usr/bin/env python
"""Time the tensor-core front-end conv of the bench workload (init_conv on T*B = 320 frames of 128x128, 8 -> 64
channels, bias + ReLU) and the cfg2 dgrad-shaped plain conv (256 -> 128 channels) with CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import plconv  # noqa: E402
from plconv import functional as F  # noqa: E402


def bench(N, H, W, cin, cout, k=3, iters=20):
    dev = torch.device("cuda:0")
    conv = torch.nn.Conv2d(cin, cout, k, padding=k // 2).to(dev)
    cp = F.ConvParams(conv, relu=True)
    x = torch.randn(N, H, W, cp.cin_p, device=dev).to(torch.bfloat16)
    out = torch.empty(N, H, W, cp.cout_p, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        F.conv2d_same_into(x, cp, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        F.conv2d_same_into(x, cp, out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    gb = (x.numel() + out.numel()) * 2 / 1e9
    print(f"conv {cin}->{cout} k{k} on {N}x{H}x{W}: {us:.1f} us  ({gb / (us * 1e-6):.0f} GB/s in+out, "
          f"{2.0 * N * H * W * cin * cout * k * k / (us * 1e-6) / 1e12:.0f} TFLOP/s)")


if __name__ == "__main__":
    bench(320, 128, 128, 8, 64)
    bench(32, 128, 128, 256, 128)
    bench(32, 128, 128, 64, 64)


def bench_fused(B=32, T=10, H=128, W=128, iters=20):
    dev = torch.device("cuda:0")
    frames = torch.rand(B, T, 1, H, W, device=dev)
    w = torch.randn(64, 3, 3, 3, device=dev) * 0.2
    b = torch.randn(64, device=dev) * 0.1
    out = torch.empty(T * B, H, W, 64, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        F.frontend_tc(frames, w, b, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        F.frontend_tc(frames, w, b, out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    gb = (frames.numel() * 4 + out.numel() * 2) / 1e9
    print(f"fused front-end (im2col in kernel) {B}x{T} frames {H}x{W} -> 64 ch: {us:.1f} us  ({gb / (us * 1e-6):.0f} GB/s "
          f"in+out)")


if __name__ == "__main__":
    bench_fused()
