#!/usr/bin/env python
"""Time one cell step (fwd, optionally bwd) with CUDA events.  python tools/microbench.py B Cin Ch H W k [--bwd] [--saved]
--saved: saved-gates form (plc_cell_fwd_save / plc_cell_bwd_saved) instead of gate recompute."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import plconv  # noqa: E402
from plconv import functional as F  # noqa: E402


def time_fn(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():  # noqa: C901
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B, cin, ch, H, W, k = (int(a) for a in args[:6]) if len(args) >= 6 else (32, 64, 64, 128, 128, 3)
    dev = torch.device("cuda:0")
    mode = plconv.PLC_MODE_FP32 if "--fp32" in sys.argv else plconv.PLC_MODE_BF16_TC
    adt = torch.float32 if "--fp32" in sys.argv else torch.bfloat16
    w = torch.randn(4 * ch, cin + ch, k, k, device=dev) * 0.02
    b = torch.zeros(4 * ch, device=dev)
    pw = F.pack_weights(w, b, cin, ch, k, mode, with_dgrad=True)
    x = torch.randn(B, H, W, cin, device=dev).to(adt) if cin else None
    h = (torch.randn(B, H, W, ch, device=dev) * 0.5).to(adt)
    c = torch.randn(B, H, W, ch, device=dev)
    h2, c2 = torch.empty_like(h), torch.empty_like(c)
    flops = 2.0 * B * H * W * (cin + ch) * k * k * 4 * ch
    saved = None
    if "--saved" in sys.argv:
        n = F.saved_gates_bytes(B, H, W, pw)
        if not n:
            raise SystemExit("this shape has no saved-gates form")
        saved = torch.empty(n, dtype=torch.uint8, device=dev)
    med, mn = time_fn(lambda: F.cell_forward(x, h, c, pw, h_out=h2, c_out=c2, saved=saved))
    print(f"fwd  B{B} {cin}->{ch} {H}x{W} k{k}: median {med:.1f} us  min {mn:.1f} us  "
          f"{flops / med / 1e6:.1f} TFLOP/s (median) {flops / mn / 1e6:.1f} (min)")
    if "--bwd" in sys.argv:
        dh = torch.randn_like(h)
        dc = torch.randn_like(c)
        dW = torch.zeros(4 * ch, cin + ch, k, k, device=dev)
        db = torch.zeros(4 * ch, device=dev)
        ws = F.bwd_workspace(B, H, W, pw, dev)
        img = F.wgrad_accumulator(B, H, W, pw, dev)
        dx, dhp, dcp = (torch.empty_like(x) if cin else None), torch.empty_like(h), torch.empty_like(c)
        med, mn = time_fn(lambda: F.cell_backward_acc(x, h, c, pw, dh, None, dc, img, db, workspace=ws, dx=dx,
                                                  dh_prev=dhp, dc_prev=dcp, saved=saved), iters=10, warm=2)
        print(f"bwd  median {med:.1f} us  min {mn:.1f} us  {2 * flops / med / 1e6:.1f} TFLOP/s (2F algorithmic)")


if __name__ == "__main__":
    main()
