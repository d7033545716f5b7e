#!/usr/bin/env python
"""Drop-in Generator train step (forward + CombinedLoss restatement + backward) at the reference's shipped shapes
(configs/default.yaml: hidden [16,32], T=5, x8, batch 8; LR 15x12 frames from test/test_loss_fix.py:47-48) and at a
larger hidden-64 setting; CUDA-event timing, plus the CPU oracle-free torch reference timing when --cpu is given.
    python tools/generator_bench.py [--cpu]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import plconv  # noqa: E402


def run(B, T, H, W, hd, scale, lu_ch, mode, iters=10):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    gen = plconv.Generator(1, 1, lu_ch, hd, scale_factor=scale, mode=mode).to(dev)
    gen.materialize(scale, dev)
    rain = torch.rand(B, T, 1, H, W, device=dev) * 5
    dem = torch.rand(B, 1, H * scale, W * scale, device=dev)
    lu = torch.rand(B, lu_ch, H * scale, W * scale, device=dev)
    n_st = 30
    s_coords = torch.stack([torch.randint(0, H, (n_st,)), torch.randint(0, W, (n_st,))], 1).to(dev)
    s_vals = (torch.rand(T, n_st) * 20).to(dev)
    opt = torch.optim.Adam(gen.parameters(), lr=5e-4)
    loss_mod = plconv.CombinedLoss()                       # fused loss + gradient (plc_combined_loss)

    def step():
        opt.zero_grad()
        pred = gen(rain, dem, lu)
        loss, _ = loss_mod(pred, rain, s_coords, s_vals, scale_factor=scale)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(gen.parameters(), 0.5)
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"Generator train step  B{B} T{T} LR {H}x{W} hidden {hd} x{scale} mode={mode}: {ms:.2f} ms/step "
          f"-> {B / ms * 1e3:.0f} sequences/s", flush=True)


def run_trainer(B, T, H, W, hd, scale, lu_ch, mode, n_batches=12, epochs=3):
    """The same step through plconv.Trainer: pinned host batches, prefetch stream, no per-step host sync
    (end to end: H2D of every batch is inside the timed region)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cfg = plconv.TrainerConfig(hidden_dims=hd, lu_channels=lu_ch, scale_factor=scale, mode=mode)
    tr = plconv.Trainer(cfg, device=dev)
    data = plconv.trainer.SyntheticRainBatches(n_batches, B, T, H, W, scale, lu_ch, n_stations=30)
    tr.train_epoch(data)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(epochs):
        out = tr.train_epoch(data)                  # ends with the epoch's single host read
    dt = (time.perf_counter() - t0) / (epochs * n_batches)
    print(f"Trainer.train_epoch   B{B} T{T} LR {H}x{W} hidden {hd} x{scale} mode={mode}: {dt * 1e3:.2f} ms/step "
          f"-> {B / dt:.0f} sequences/s (host wall clock incl. H2D; loss {out['total']:.3f})", flush=True)


if __name__ == "__main__":
    run_trainer(8, 5, 15, 12, [16, 32], 8, 5, "bf16")
    run_trainer(16, 10, 64, 64, [64, 64], 4, 5, "bf16")
    for mode in ("bf16", "fp32"):
        run(8, 5, 15, 12, [16, 32], 8, 5, mode)          # shipped default shapes (reference CPU: 415 ms/step, 19.3 seq/s)
    run(4, 10, 64, 64, [16, 32], 1, 5, "bf16")           # BASELINE cfg-1 shapes (reference CPU: 388 ms/step)
    run(16, 10, 64, 64, [64, 64], 4, 5, "bf16")          # hidden 64, x4 upsampling


def run_infer(B, T, H, W, hd, scale, lu_ch, mode="bf16", iters=50):
    """Inference forward: eager vs CUDA-graph replay (Generator.forward_graphed)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    gen = plconv.Generator(1, 1, lu_ch, hd, scale_factor=scale, mode=mode).to(dev).eval()
    gen.materialize(scale, dev)
    a = (torch.rand(B, T, 1, H, W, device=dev) * 5, torch.rand(B, 1, H * scale, W * scale, device=dev),
         torch.rand(B, lu_ch, H * scale, W * scale, device=dev))

    def timeit(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e3

    with torch.no_grad():
        te = timeit(lambda: gen(*a))
    tg = timeit(lambda: gen.forward_graphed(*a))
    print(f"Generator inference   B{B} T{T} LR {H}x{W} hidden {hd} x{scale} mode={mode}: eager {te:.2f} ms, "
          f"graph replay {tg:.2f} ms -> {B / tg * 1e3:.0f} sequences/s", flush=True)


if __name__ == "__main__":
    run_infer(8, 5, 15, 12, [16, 32], 8, 5)
    run_infer(4, 10, 64, 64, [16, 32], 1, 5)
