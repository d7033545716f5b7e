#!/usr/bin/env python
"""In-kernel cycle accounting of the fused cell-step kernel (plc_debug_set_prof): where does the MMA warp wait?
    python tools/kprof.py B Cin Ch H W k"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_KLIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "ubench", "libplc_kprof.so")
if not os.path.exists(_KLIB):
    raise SystemExit("build the instrumented library first:  python pl-convlstm-gan_b200/build.py --kprof")
os.environ.setdefault("PLC_LIB", _KLIB)   # cycle counters only exist in the -DPLC_KPROF build
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import plconv  # noqa: E402
from plconv import functional as F  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [32, 64, 64, 128, 128, 3]
    B, cin, ch, H, W, k = a
    dev = torch.device("cuda:0")
    lib = plconv._lib.load()
    w = torch.randn(4 * ch, cin + ch, k, k, device=dev) * 0.02
    pw = F.pack_weights(w, torch.zeros(4 * ch, device=dev), cin, ch, k, plconv.PLC_MODE_BF16_TC)
    x = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
    h = torch.randn(B, H, W, ch, device=dev).to(torch.bfloat16)
    c = torch.randn(B, H, W, ch, device=dev)
    h2, c2 = torch.empty_like(h), torch.empty_like(c)
    for _ in range(3):
        F.cell_forward(x, h, c, pw, h_out=h2, c_out=c2)
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    lib.plc_debug_set_prof(ctypes.c_void_p(buf.data_ptr()))
    F.cell_forward(x, h, c, pw, h_out=h2, c_out=c2)
    torch.cuda.synchronize()
    lib.plc_debug_set_prof(None)
    p = buf.view(148, 16).cpu().double()
    lead = p[p[:, 0] > 0]
    print(f"leader CTAs: {len(lead)}")
    tot, te, tf, tiles = lead[:, 0].mean(), lead[:, 1].mean(), lead[:, 2].mean(), lead[:, 3].mean()
    print(f"MMA warp: total {tot:.0f} cyc, tiles {tiles:.1f}, per tile {tot / tiles:.0f}")
    print(f"  wait TMEM-empty (epilogue) {te:.0f} ({100 * te / tot:.1f}%)   wait TMA-full {tf:.0f} ({100 * tf / tot:.1f}%)"
          f"   issuing {tot - te - tf:.0f} ({100 * (tot - te - tf) / tot:.1f}%)")
    ep = p[p[:, 5] > 0]
    t = max(tiles, 1)
    print(f"  per tile: barrier-A {ep[:, 6].mean() / t:.0f}  tmem-ld-wait {ep[:, 7].mean() / t:.0f}  math+st.shared "
          f"{ep[:, 8].mean() / t:.0f}  fence+barrier-B {ep[:, 9].mean() / t:.0f}  decode+prefetch {ep[:, 10].mean() / t:.0f}  "
          f"store-issue {ep[:, 11].mean() / t:.0f}")
    print(f"epilogue warp 4 (all CTAs {len(ep)}): idle-wait {ep[:, 4].mean():.0f} cyc, busy {ep[:, 5].mean():.0f} cyc, "
          f"busy per tile {ep[:, 5].mean() / max(tiles - 1, 1):.0f}")


if __name__ == "__main__":
    main()
