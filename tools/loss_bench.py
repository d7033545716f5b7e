#!/usr/bin/env python
"""Time plc_combined_loss (fwd + d total/d pred) against the eager restatement of the reference loss on the same GPU.

Algorithmic bytes per HR pixel: read pred once in pass A, once in pass B, write dpred once = 12 B
(+ 8 B / s^2 for the LR grid), so achieved GB/s = 12 * B*T*Hs*Ws / time against the measured HBM copy peak.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plconv  # noqa: E402


def eager_combined_loss(pred, lr, coords, obs, scale, lambdas=(1.0, 1.0, 0.1, 0.05)):
    """The reference's loss as plain eager torch ops (combined_loss.py:64-191, log weights) -- the comparison arm of
    this benchmark (written out here: only tests/ and bench.py may use oracle/)."""
    import torch.nn.functional as F
    b, t, c, h, w = pred.shape
    hl, wl = lr.shape[-2:]
    pooled = F.interpolate(pred.reshape(b * t, c, h, w), size=(hl, wl), mode="area").reshape(b, t, c, hl, wl)
    cons = (pooled - lr).abs().mean()
    sc = ((coords.float() + 0.5) * scale - 0.5).long()
    ok = (sc[:, 0] >= 0) & (sc[:, 0] < h) & (sc[:, 1] >= 0) & (sc[:, 1] < w)
    at = pred[:, :, 0][:, :, sc[ok, 0], sc[ok, 1]]
    o = obs[:, ok].unsqueeze(0).expand(b, -1, -1)
    m = ~torch.isnan(o)
    point = ((at[m] - o[m]).abs() * (1.0 + torch.log1p(o[m]))).mean()
    smooth = (pred[..., :, :-1] - pred[..., :, 1:]).abs().mean() + (pred[..., :-1, :] - pred[..., 1:, :]).abs().mean()
    temp = (pred[:, :-1] - pred[:, 1:]).abs().mean()
    lp, lc, ls, lt = lambdas
    return lp * point + lc * cons + ls * smooth + lt * temp


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"]
    mod = plconv.CombinedLoss()
    cases = [(8, 5, 30, 24, 4, 60), (64, 20, 128, 128, 1, 300), (32, 10, 128, 128, 4, 300)]
    pick = [int(a) for a in sys.argv[1:]] or range(len(cases))       # e.g. `loss_bench.py 2` under ncu
    for (B, T, H, W, s, n_st) in [cases[i] for i in pick]:
        pred = (torch.rand(B, T, 1, H * s, W * s, device=dev) * 10).requires_grad_(True)
        lr = torch.rand(B, T, 1, H, W, device=dev) * 10
        coords = torch.stack([torch.randint(0, H, (n_st,)), torch.randint(0, W, (n_st,))], 1).to(dev)
        obs = (torch.rand(T, n_st) * 40).to(dev)

        def fused():
            pred.grad = None
            total, _ = mod(pred, lr, coords, obs, scale_factor=s)
            total.backward()

        def eager():
            pred.grad = None
            total = eager_combined_loss(pred, lr, coords, obs, s)
            total.backward()

        tf, te = timeit(fused), timeit(eager)
        npx = B * T * H * s * W * s
        gbs = 12.0 * npx / (tf * 1e-6) / 1e9
        print(json.dumps({"shape": [B, T, H, W, s], "hr_pixels": npx, "fused_us": round(tf, 1), "eager_us": round(te, 1),
                          "speedup": round(te / tf, 2), "algorithmic_GBps": round(gbs, 1),
                          "frac_of_hbm_peak": round(gbs / hbm, 3)}))


if __name__ == "__main__":
    main()
