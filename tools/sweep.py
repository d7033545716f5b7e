#!/usr/bin/env python
"""BASELINE.json configs[4]: ConvLSTM cell microbench sweep (hidden 32-256, kernel 3/5, H/W 64-512), forward and
BPTT cell step, CUDA events (median of 10 after 3 warm-ups), fraction of the measured bf16 peak -- next to the
REFERENCE cell (convlstm.py:16-28 restated inline: cat -> conv2d -> split -> sigmoid/tanh -> update) run eagerly on the
same GPU (fp32 parameters, torch's default flags = cuDNN with TF32 convolutions allowed, autograd backward) and on this
box's CPU cores (one sample, scaled linearly to the batch; only where a step stays under ~2 s).  Reported baselines only.
Writes a markdown table to stdout.   python tools/sweep.py [--quick] [--no-cpu]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import plconv  # noqa: E402
from plconv import functional as F  # noqa: E402
from microbench import time_fn  # noqa: E402


def ref_cell(x, h, c, w, b):
    """The reference's ConvLSTMCell.forward (src/models/convlstm.py:16-28), NCHW fp32."""
    k = w.shape[-1]
    z = torch.nn.functional.conv2d(torch.cat([x, h], 1), w, b, padding=k // 2)
    i, f, o, g = torch.split(z, h.shape[1], 1)
    i, f, o, g = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o), torch.tanh(g)
    c2 = f * c + i * g
    return o * torch.tanh(c2), c2


def ref_times(B, ch, hw, k, dev, iters=5, warm=2):
    """(fwd us, fwd+bwd us) of the eager reference cell on `dev`."""
    import time
    w = (torch.randn(4 * ch, 2 * ch, k, k, device=dev) * 0.02).requires_grad_()
    b = torch.zeros(4 * ch, device=dev, requires_grad=True)
    x = torch.randn(B, ch, hw, hw, device=dev, requires_grad=True)
    h = (torch.randn(B, ch, hw, hw, device=dev) * 0.5).requires_grad_()
    c = torch.randn(B, ch, hw, hw, device=dev, requires_grad=True)
    gh, gc = torch.randn(B, ch, hw, hw, device=dev), torch.randn(B, ch, hw, hw, device=dev)

    def fwd():
        with torch.no_grad():
            ref_cell(x, h, c, w, b)

    def both():
        for t in (w, b, x, h, c):
            t.grad = None
        h2, c2 = ref_cell(x, h, c, w, b)
        torch.autograd.backward([h2, c2], [gh, gc])

    if dev.type == "cuda":
        return time_fn(fwd, iters=iters, warm=warm)[0], time_fn(both, iters=iters, warm=warm)[0]
    out = []
    for fn in (fwd, both):
        fn()
        t0 = time.perf_counter()
        for _ in range(2):
            fn()
        out.append((time.perf_counter() - t0) / 2 * 1e6)
    return tuple(out)


def main():
    quick = "--quick" in sys.argv
    dev = torch.device("cuda:0")
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["bf16_tflops_sustained"] if os.path.exists(peaks) else 1590.0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    print(f"reference columns: eager convlstm.py cell, fp32; GPU = same B200 (cuDNN, TF32 convs allowed = torch default), "
          f"CPU = {cores} cores, 1 sample x B (linear)\n")
    print(f"| B | Cin=Ch | HxW | k | fwd us | fwd TFLOP/s | frac of {peak:.0f} | bwd us | fwd+bwd TFLOP/s (3F) | frac | "
          f"fwd+bwd us saved gates | ref GPU fwd us | ref GPU fwd+bwd us | speed-up fwd / fwd+bwd | ref CPU fwd+bwd ms |")
    print("|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|")
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
                sv_s = "-"                                          # saved-gates form (Ch % 64 == 0): fwd_save + bwd_saved
                nsv = F.saved_gates_bytes(B, hw, hw, pw)
                if nsv:
                    saved = torch.empty(nsv, dtype=torch.uint8, device=dev)
                    fs, _ = time_fn(lambda: F.cell_forward(x, h, c, pw, h_out=h2, c_out=c2, saved=saved), iters=10, warm=3)
                    bs, _ = time_fn(lambda: F.cell_backward_acc(x, h, c, pw, dh, None, dc, img, db, workspace=ws, dx=dx,
                                                                dh_prev=dhp, dc_prev=dcp, saved=saved), iters=10, warm=3)
                    sv_s = f"{fs + bs:.1f}"
                    del saved
                tf_f = flops / fwd / 1e6
                tf_t = 3 * flops / (fwd + bwd) / 1e6
                del x, h, c, h2, c2, dh, dc, ws, dx, dhp, dcp, img
                torch.cuda.empty_cache()
                try:
                    rf, rb = ref_times(B, ch, hw, k, dev)
                    ref_s = f"{rf:.0f} | {rb:.0f} | {rf / fwd:.1f}x / {rb / (fwd + bwd):.1f}x"
                except torch.OutOfMemoryError:
                    ref_s = "OOM | OOM | -"
                torch.cuda.empty_cache()
                cpu_s = "-"
                if "--no-cpu" not in sys.argv and 3 * flops / B < 0.6e12:      # ~2 s per step at ~0.3 TFLOP/s
                    _, cb = ref_times(1, ch, hw, k, torch.device("cpu"))
                    cpu_s = f"{cb * B / 1e3:.0f}"
                print(f"| {B} | {ch} | {hw}x{hw} | {k} | {fwd:.1f} | {tf_f:.0f} | {tf_f / peak:.2f} | {bwd:.1f} | "
                      f"{tf_t:.0f} | {tf_t / peak:.2f} | {sv_s} | {ref_s} | {cpu_s} |", flush=True)


if __name__ == "__main__":
    main()
