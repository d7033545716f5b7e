#!/usr/bin/env python
"""Cycle accounting of the gate-recompute kernel alone (plc_cell_bwd with no dgrad / wgrad outputs)."""
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
from plconv._lib import PlcCellDesc

_a = [v for v in sys.argv[1:] if not v.startswith("--")]
a = [int(v) for v in _a[:6]] if len(_a) >= 6 else [32, 64, 64, 128, 128, 3]
B, cin, ch, H, W, k = a
dev = torch.device("cuda:0")
lib = plconv._lib.load()
w = torch.randn(4 * ch, cin + ch, k, k, device=dev) * 0.02
pw = F.pack_weights(w, torch.zeros(4 * ch, device=dev), cin, ch, k, plconv.PLC_MODE_BF16_TC, with_dgrad=True)
x = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
h = torch.randn(B, H, W, ch, device=dev).to(torch.bfloat16)
c = torch.randn(B, H, W, ch, device=dev)
dh, dh2, dc = torch.randn_like(h), torch.randn_like(h), torch.randn_like(c)
dcp = torch.empty_like(c)
ws = F.bwd_workspace(B, H, W, pw, dev)
d = PlcCellDesc(B, H, W, cin, ch, k, 0, 1)
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def run():
    rc = lib.plc_cell_bwd(ctypes.byref(d), P(x), P(h), P(c), P(pw.fwd), P(pw.dgrad), P(pw.bias), P(dh), P(dh2), P(dc),
                          None, None, P(dcp), None, None, P(ws), ws.numel(), st)
    assert rc == 0, lib.plc_last_error()

for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(10):
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f"gates kernel {ts[len(ts) // 2]:.1f} us median, {ts[0]:.1f} us min  ({os.environ.get('PLC_LIB')})")
if "--time-only" in sys.argv:
    sys.exit(0)
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.plc_debug_set_prof(ctypes.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
lib.plc_debug_set_prof(None)
p = buf.view(148, 16).cpu().double()
lead = p[p[:, 0] > 0]
tot, te, tf, tiles = lead[:, 0].mean(), lead[:, 1].mean(), lead[:, 2].mean(), lead[:, 3].mean()
print(f"MMA warp per tile {tot / tiles:.0f} cyc; wait TMEM-empty {100 * te / tot:.1f}%  wait TMA-full {100 * tf / tot:.1f}%")
ep = p[p[:, 5] > 0]
print(f"epilogue warp 4: idle {ep[:, 4].mean():.0f}, busy per tile {ep[:, 5].mean() / max(tiles - 1, 1):.0f}")
t = max(tiles, 1)
print(f"  per tile: operand wait {ep[:, 6].mean() / t:.0f}  tmem-ld-wait {ep[:, 7].mean() / t:.0f}  math+ld/st.shared {ep[:, 8].mean() / t:.0f}  "
      f"barrier X {ep[:, 9].mean() / t:.0f}  decode {ep[:, 10].mean() / t:.0f}")
