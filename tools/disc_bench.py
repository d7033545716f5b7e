#!/usr/bin/env python
"""Per-launch timing of one discriminator training pass (forward + backward with dX and dW) at the cfg3 shapes:
    python tools/disc_bench.py [N=128] [T=20] [H=128] [W=128]
Prints every library launch in order (kind, microseconds, executed TFLOP/s) through plc_timing_*."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import plconv  # noqa: E402
from plconv import _lib  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:5]]
    n, t, hh, ww = (a + [128, 20, 128, 128][len(a):])
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    disc = plconv.Discriminator().to(dev)
    clip = torch.relu(torch.randn(n, t, 1, hh, ww, device=dev) + 0.3).requires_grad_()

    def one():
        for p in disc.parameters():
            p.grad = None
        clip.grad = None
        plconv.gan.bce_with_logits(disc(clip), 1.0).backward()

    for _ in range(3):
        one()
    torch.cuda.synchronize()
    _lib.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    one()
    e1.record()
    torch.cuda.synchronize()
    rec = _lib.timing_collect()
    _lib.timing_enable(False)
    tot = 0.0
    for kind, ms, fl in rec:
        tot += ms
        print(f"{kind:12s} {ms * 1e3:9.1f} us  {fl / 1e9:9.2f} GFLOP  {fl / (ms * 1e-3) / 1e12 if fl else 0:8.1f} TFLOP/s")
    print(f"library launches {len(rec)}: {tot:.3f} ms of {e0.elapsed_time(e1):.3f} ms (fwd + bwd, one clip batch of {n})")


if __name__ == "__main__":
    main()
