#!/usr/bin/env python
"""In-kernel cycle accounting of a plain tensor-core conv (plc_conv_fwd), e.g. the frame front-end 8 -> 64:
    python tools/kprof_conv.py N H W Cin Cout k"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_KLIB = os.path.join(ROOT, "tools", "ubench", "libplc_kprof.so")
if not os.path.exists(_KLIB):
    raise SystemExit("build the instrumented library first:  python pl-convlstm-gan_b200/build.py --kprof")
os.environ.setdefault("PLC_LIB", _KLIB)
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import plconv  # noqa: E402
from plconv import functional as F  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [320, 128, 128, 8, 64, 3]
    N, H, W, cin, cout, k = a
    dev = torch.device("cuda:0")
    lib = plconv._lib.load()
    conv = torch.nn.Conv2d(cin, cout, k, padding=k // 2).to(dev)
    cp = F.ConvParams(conv, relu=True)
    x = torch.randn(N, H, W, cp.cin_p, device=dev).to(torch.bfloat16)
    out = torch.empty(N, H, W, cp.cout_p, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        F.conv2d_same_into(x, cp, out)
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    lib.plc_debug_set_prof(ctypes.c_void_p(buf.data_ptr()))
    F.conv2d_same_into(x, cp, out)
    torch.cuda.synchronize()
    lib.plc_debug_set_prof(None)
    p = buf.view(148, 16).cpu().double()
    lead = p[p[:, 0] > 0]
    tot, te, tf, tiles, tp = (lead[:, i].mean() for i in (0, 1, 2, 3, 13))
    print(f"leader CTAs {len(lead)}; MMA warp: total {tot:.0f} cyc, tiles {tiles:.1f}, per tile {tot / tiles:.0f}")
    print(f"  per tile: wait TMEM-empty {te / tiles:.0f}  wait patch {tp / tiles:.0f}  wait weights {tf / tiles:.0f}  "
          f"issue {(tot - te - tf - tp) / tiles:.0f} (of which MMA loop {lead[:, 14].mean() / tiles:.0f})")
    ep = p[p[:, 5] > 0]
    print(f"epilogue warp 4: idle-wait per tile {ep[:, 4].mean() / tiles:.0f}, busy per tile {ep[:, 5].mean() / max(tiles - 1, 1):.0f}")


if __name__ == "__main__":
    main()
