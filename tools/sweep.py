#!/usr/bin/env python
"""BASELINE.json configs[4]: ConvLSTM cell microbench sweep (hidden 32-256, kernel 3/5, H/W 64-512), forward and
BPTT cell step, CUDA events (median of 10 after 3 warm-ups), fraction of the measured bf16 peak.
Writes a markdown table to stdout.   python tools/sweep.py [--quick]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import plconv  # noqa: E402
from plconv import functional as F  # noqa: E402
from microbench import time_fn  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    dev = torch.device("cuda:0")
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["bf16_tflops_sustained"] if os.path.exists(peaks) else 1590.0
    print(f"| B | Cin=Ch | HxW | k | fwd us | fwd TFLOP/s | frac of {peak:.0f} | bwd us | fwd+bwd TFLOP/s (3F) | frac |")
    print("|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|")
    chs = [32, 64, 128, 256]
    sizes = [64, 128, 256] if quick else [64, 128, 256, 512]
    for ch in chs:
        for k in (3, 5):
            for hw in sizes:
                B = 8
                m = B * hw * hw
                act_bytes = m * ch * (2 + 2 + 4 + 4 + 2 + 4 * 2 * 2) * 3
                if act_bytes > 60e9:
                    continue
                flops = 2.0 * m * (2 * ch) * k * k * 4 * ch
                w = torch.randn(4 * ch, 2 * ch, k, k, device=dev) * 0.02
                b = torch.zeros(4 * ch, device=dev)
                pw = F.pack_weights(w, b, ch, ch, k, plconv.PLC_MODE_BF16_TC, with_dgrad=True)
                x = torch.randn(B, hw, hw, ch, device=dev).to(torch.bfloat16)
                h = (torch.randn(B, hw, hw, ch, device=dev) * 0.5).to(torch.bfloat16)
                c = torch.randn(B, hw, hw, ch, device=dev)
                h2, c2 = torch.empty_like(h), torch.empty_like(c)
                fwd, _ = time_fn(lambda: F.cell_forward(x, h, c, pw, h_out=h2, c_out=c2), iters=10, warm=3)
                dh, dc = torch.randn_like(h), torch.randn_like(c)
                dW = torch.zeros(4 * ch, 2 * ch, k, k, device=dev)
                db = torch.zeros(4 * ch, device=dev)
                ws = F.bwd_workspace(B, hw, hw, pw, dev)
                img = F.wgrad_accumulator(B, hw, hw, pw, dev)
                dx, dhp, dcp = torch.empty_like(x), torch.empty_like(h), torch.empty_like(c)
                bwd, _ = time_fn(lambda: F.cell_backward_acc(x, h, c, pw, dh, None, dc, img, db, workspace=ws, dx=dx,
                                                         dh_prev=dhp, dc_prev=dcp), iters=10, warm=3)
                tf_f = flops / fwd / 1e6
                tf_t = 3 * flops / (fwd + bwd) / 1e6
                print(f"| {B} | {ch} | {hw}x{hw} | {k} | {fwd:.1f} | {tf_f:.0f} | {tf_f / peak:.2f} | {bwd:.1f} | "
                      f"{tf_t:.0f} | {tf_t / peak:.2f} |", flush=True)
                del x, h, c, h2, c2, dh, dc, ws, dx, dhp, dcp
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
