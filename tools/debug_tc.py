#!/usr/bin/env python
"""Run each tensor-core forward case in its own subprocess (a hang or fault kills only that case) and
print an error summary.  Developer tool for the GPU box:  python tools/debug_tc.py [case ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (B, Cin, Ch, H, W, k)
    "gemm_k1": (1, 64, 64, 8, 16, 1),
    "k3_onetile": (1, 64, 64, 8, 16, 3),
    "k3_multi": (2, 64, 64, 16, 32, 3),
    "c16": (1, 16, 16, 12, 15, 3),
    "ch128": (1, 128, 128, 16, 16, 3),
    "cfg2_small": (4, 64, 64, 128, 128, 3),
}


def run_case(name):
    import torch
    from oracle import convlstm_oracle as O
    import plconv
    from plconv import functional as F
    from test_cell_gpu import bf16r, nhwc, nchw, pad8, report
    B, cin, ch, H, W, k = CASES[name]
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(1)
    fan_in = (cin + ch) * k * k
    w = (torch.rand(4 * ch, cin + ch, k, k, generator=gen) * 2 - 1) * (3.0 / fan_in) ** 0.5 * 2
    b = torch.randn(4 * ch, generator=gen) * 0.5
    x = torch.randn(B, cin, H, W, generator=gen)
    h = torch.randn(B, ch, H, W, generator=gen) * 0.5
    c = torch.randn(B, ch, H, W, generator=gen)
    pw = F.pack_weights(w.to(dev), b.to(dev), cin, ch, k, plconv.PLC_MODE_BF16_TC, cin_pad=pad8(cin))
    xd, hd, cd = nhwc(x, torch.bfloat16, dev, pad8(cin)), nhwc(h, torch.bfloat16, dev), nhwc(c, torch.float32, dev)
    gates = torch.zeros(B, H, W, 4 * ch, dtype=torch.bfloat16, device=dev)
    h2, c2 = F.cell_forward(xd, hd, cd, pw, gates_out=gates)
    torch.cuda.synchronize()
    h_ref, c_ref, (gi, gf, go, gg) = O.cell_forward_gates(bf16r(x).double(), bf16r(h).double(), c.double(),
                                                          bf16r(w).double(), b.double())
    print(name, report("h", nchw(h2), h_ref))
    print(name, report("c", nchw(c2), c_ref))
    gref = torch.cat([gi, gf, go, gg], dim=1)
    print(name, report("gates", nchw(gates), gref))
    for gname, lo in (("i", 0), ("f", ch), ("o", 2 * ch), ("g", 3 * ch)):
        print(name, report("gate_" + gname, nchw(gates)[:, lo:lo + ch], gref[:, lo:lo + ch]))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True,
                               text=True, timeout=int(os.environ.get("PLC_CASE_TIMEOUT", "120")))
            print(f"--- {n}: exit {r.returncode}")
            print(r.stdout[-3000:])
            if r.returncode:
                print(r.stderr[-3000:])
        except subprocess.TimeoutExpired:
            print(f"--- {n}: TIMEOUT (hang)")
        sys.stdout.flush()
