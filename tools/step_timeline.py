#!/usr/bin/env python
"""GPU timeline of ONE cfg3 GAN training step replayed as a CUDA graph (torch.profiler / CUPTI kernel records):
where the step's wall time goes -- kernels by name, and the idle gaps between consecutive kernels.

    python tools/step_timeline.py [B=64] [--eager]
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import plconv  # noqa: E402
from plconv.gan import Discriminator, GanTrainStep  # noqa: E402
from plconv.training import GraphedStep  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = int(args[0]) if args else 64
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    gen = plconv.NowcastGenerator(1, [64, 64], 3, 10, 10, "bf16").to(dev)
    disc = Discriminator().to(dev)
    step = GanTrainStep(gen, disc, lr_g=5e-4, lr_d=2e-4, lambda_adv=0.05, grad_clip_norm=0.5)
    frames = torch.relu(torch.randn(B, 10, 1, 128, 128, device=dev) + 0.3)
    target = torch.relu(torch.randn(B, 10, 1, 128, 128, device=dev) + 0.3)
    run = step if "--eager" in sys.argv else GraphedStep(step, (frames, target), warmup=3)
    for _ in range(3):
        run(frames, target)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        run(frames, target)
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
            ev.append((e.time_range.start, e.time_range.end, e.name))
    ev.sort()
    if not ev:
        raise SystemExit("no CUDA kernel records (CUPTI unavailable?)")
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    busy, cur_end = 0.0, t0
    gaps = []
    by_name = collections.OrderedDict()
    prev = None
    for s, e, n in ev:
        d = by_name.setdefault(n[:90], [0, 0.0])
        d[0] += 1
        d[1] += e - s
        if s > cur_end:
            gaps.append((s - cur_end, prev, n))
        if e > cur_end:
            busy += e - max(s, cur_end)
            cur_end = e
        prev = n
    span = t1 - t0
    print(f"span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {100 * (span - busy) / span:.1f}% in {len(gaps)} gaps; "
          f"{len(ev)} GPU activities")
    print("\n| n | total ms | avg us | share of span | kernel |\n|---:|---:|---:|---:|---|")
    for n, (c, t) in sorted(by_name.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"| {c} | {t / 1e3:.2f} | {t / c:.1f} | {100 * t / span:.1f}% | `{n}` |")
    gaps.sort(reverse=True)
    print(f"\ngap histogram: >50us {sum(g[0] > 50 for g in gaps)}, 10-50us {sum(10 < g[0] <= 50 for g in gaps)}, "
          f"3-10us {sum(3 < g[0] <= 10 for g in gaps)}, <=3us {sum(g[0] <= 3 for g in gaps)}; "
          f"sum of gaps >10us: {sum(g[0] for g in gaps if g[0] > 10) / 1e3:.2f} ms")
    print("\nlargest gaps (us, after -> before):")
    for g, a, b in gaps[:25]:
        print(f"  {g:8.1f}  {str(a)[:60]} -> {b[:60]}")


if __name__ == "__main__":
    main()
