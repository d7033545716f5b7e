#!/usr/bin/env python
"""In-kernel cycle accounting of the pair wgrad kernel.  python tools/kprof_wgrad.py B C Ch H W k"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_KLIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "ubench", "libplc_kprof.so")
if not os.path.exists(_KLIB):
    raise SystemExit("build the instrumented library first:  python pl-convlstm-gan_b200/build.py --kprof")
os.environ.setdefault("PLC_LIB", _KLIB)   # cycle counters only exist in the -DPLC_KPROF build
sys.path.insert(0, ROOT)
import torch
import plconv
from plconv import functional as F

a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [32, 64, 64, 128, 128, 3]
B, cin, ch, H, W, k = a
dev = torch.device("cuda:0")
lib = plconv._lib.load()
w = torch.randn(4 * ch, cin + ch, k, k, device=dev) * 0.02
pw = F.pack_weights(w, torch.zeros(4 * ch, device=dev), cin, ch, k, plconv.PLC_MODE_BF16_TC, with_dgrad=True)
x = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
h = torch.randn(B, H, W, ch, device=dev).to(torch.bfloat16)
c = torch.randn(B, H, W, ch, device=dev)
dh, dc = torch.randn_like(h), torch.randn_like(c)
img = F.wgrad_accumulator(B, H, W, pw, dev)
db = torch.zeros(4 * ch, device=dev)
ws = F.bwd_workspace(B, H, W, pw, dev)
for _ in range(2):
    F.cell_backward_acc(x, h, c, pw, dh, None, dc, img, db, workspace=ws)
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.plc_debug_set_prof(ctypes.c_void_p(buf.data_ptr()))
F.cell_backward_acc(x, h, c, pw, dh, None, dc, img, db, workspace=ws)
torch.cuda.synchronize()
lib.plc_debug_set_prof(None)
p = buf.view(148, 16).cpu().double()
# the last kernel writing the buffer is wgrad (slots 0,2,3 by leaders; 4,5 by all CTAs)
lead = p[p[:, 3] > 0]
print(f"wgrad leaders {len(lead)}: MMA warp total {lead[:,0].mean():.0f} cyc, blocks {lead[:,3].mean():.1f}, "
      f"per block {lead[:,0].mean()/lead[:,3].mean():.0f}, wait-full {100*lead[:,2].mean()/lead[:,0].mean():.1f}%")
ep = p[p[:, 5] > 0]
print(f"epilogue: wait-acc {ep[:,4].mean():.0f} cyc, flush {ep[:,5].mean():.0f} cyc")
# per output-column group (pair index = blockIdx / 2; s = pair % S, group = (pair / S) % num_groups); group 0 also
# carries the bias-gradient MMA.  PLC_WGRAD_GB=4 forces 4-block groups (N = 256 MMAs only, last group N = 128).
CB = k * k * ((cin + 63) // 64 + (ch + 63) // 64)
GB = int(os.environ.get("PLC_WGRAD_GB", "0")) or 2 * ((-(-CB // (-(-CB // 6))) + 1) // 2)
groups = -(-CB // GB)
S = max(1, 74 // groups)
print(f"CB {CB} GB {GB} groups {groups} S {S}")
idx = torch.arange(148)[p[:, 3] > 0] // 2
for g in range(groups):
    sel = lead[(idx // S) % groups == g]
    if len(sel):
        print(f"  group {g}: {len(sel)} pairs, MMA warp total {sel[:,0].mean():.0f} cyc (min {sel[:,0].min():.0f}, max "
              f"{sel[:,0].max():.0f}), per block {sel[:,0].mean()/sel[:,3].mean():.0f}, wait-full "
              f"{100*sel[:,2].mean()/sel[:,0].mean():.1f}%")
