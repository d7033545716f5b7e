"""Does an HBM-bound kernel overlap with the tensor-bound dgrad / wgrad kernels when both are resident?  (power / co-residency probe)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plconv
from plconv import functional as F
dev = torch.device("cuda:0")
B, cin, ch, H, W, k = 64, 64, 64, 128, 128, 3
w = torch.randn(4 * ch, cin + ch, k, k, device=dev) * 0.02
pw = F.pack_weights(w, torch.zeros(4 * ch, device=dev), cin, ch, k, plconv.PLC_MODE_BF16_TC, with_dgrad=True)
x = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
h = (torch.randn(B, H, W, ch, device=dev) * 0.5).to(torch.bfloat16)
c = torch.randn(B, H, W, ch, device=dev)
dh, dc = torch.randn_like(h), torch.randn_like(c)
img = F.wgrad_accumulator(B, H, W, pw, dev); db = torch.zeros(4 * ch, device=dev)
ws = F.bwd_workspace(B, H, W, pw, dev)
saved = torch.empty(F.saved_gates_bytes(B, H, W, pw), dtype=torch.uint8, device=dev)
h2, c2 = torch.empty_like(h), torch.empty_like(c)
F.cell_forward(x, h, c, pw, h_out=h2, c_out=c2, saved=saved)
dx, dhp, dcp = torch.empty_like(x), torch.empty_like(h), torch.empty_like(c)
big_a = torch.randn(256 * 1024 * 1024, device=dev)     # 1 GiB fp32
big_b = torch.empty_like(big_a)
side = torch.cuda.Stream()

def tensor_work(n):      # gates(saved: HBM) + dgrad + wgrad (tensor) per call
    for _ in range(n):
        F.cell_backward_acc(x, h, c, pw, dh, None, dc, img, db, workspace=ws, dx=dx, dh_prev=dhp, dc_prev=dcp, saved=saved)

def fwd_work(n):
    for _ in range(n):
        F.cell_forward(x, h, c, pw, h_out=h2, c_out=c2)

def hbm_work(n):         # 2 GiB of traffic per copy
    for _ in range(n):
        big_b.copy_(big_a)

def timeit(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)

for name, work in (("bwd call (gates-saved + dgrad + wgrad)", tensor_work), ("fwd", fwd_work)):
    work(5); hbm_work(5)
    n = 40
    t_t = timeit(lambda: work(n))
    t_h = timeit(lambda: hbm_work(n))
    def both():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            hbm_work(n)
        work(n)
        torch.cuda.current_stream().wait_stream(side)
    t_b = timeit(both)
    print(f"{name}: tensor alone {t_t:.1f} ms, copy alone {t_h:.1f} ms ({n * 2.147 / t_h:.2f} TB/s), together {t_b:.1f} ms "
          f"(sum {t_t + t_h:.1f}, max {max(t_t, t_h):.1f})")
