#!/usr/bin/env python
"""Eager launches vs CUDA-graph replay of the inference rollout on launch-bound shapes.  python tools/graph_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import plconv  # noqa: E402


def bench(B, H, W, hd, t_in, t_out):
    dev = torch.device("cuda:0")
    model = plconv.NowcastGenerator(1, hd, 3, t_in, t_out, "bf16").to(dev)
    runner = plconv.NowcastRunner(model, B, H, W, dev)
    frames = torch.rand(B, t_in, 1, H, W, device=dev)

    def timeit(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    eager = timeit(lambda: runner.run(frames))
    runner.capture(frames)
    graph = timeit(lambda: runner.replay(frames))
    print(f"B{B} {H}x{W} hidden {hd} T={t_in}->{t_out}: eager {eager * 1e3:.0f} us/rollout, graph {graph * 1e3:.0f} us/rollout "
          f"({eager / graph:.2f}x), {B / graph * 1e3:.0f} sequences/s", flush=True)


if __name__ == "__main__":
    bench(4, 64, 64, [16, 32], 10, 10)      # BASELINE cfg-1 shapes
    bench(8, 16, 16, [16, 32], 5, 5)        # shipped-default-like LR frames
    bench(8, 128, 128, [64, 64], 10, 10)    # cfg3 @ 8 GPUs per-GPU shard
    bench(32, 128, 128, [64, 64], 10, 10)   # cfg2
