#!/usr/bin/env python
"""Chunked weight gradient vs one wgrad per step at the bench's shapes (cfg3 generator, 8 and 64 sequences): largest
relative deviation of every parameter gradient, next to the run-to-run noise of the per-step form (red.add order).
    python tools/check_defer_wgrad.py        (PLC_DEFER_WGRAD semantics: plconv.nn.DEFER_WGRAD = "off" | "auto" | "<steps>")"""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plconv
from plconv import nn as pnn
torch.manual_seed(5)
dev = torch.device("cuda:0")
for B in (8, 64):
    model = plconv.NowcastGenerator(1, [64, 64], 3, 10, 10, "bf16").to(dev)
    frames = torch.relu(torch.randn(B, 10, 1, 128, 128, device=dev) + 0.3)
    tgt = torch.relu(torch.randn(B, 10, 1, 128, 128, device=dev) + 0.3)
    def run(m):
        pnn.DEFER_WGRAD = m
        for p in model.parameters():
            p.grad = None
        ((model(frames) - tgt).abs().mean()).backward()
        torch.cuda.synchronize()
        return {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    a, b, c = run("off"), run("auto"), run("off")
    err = lambda x, y: float((x.double() - y.double()).abs().max() / (y.double().abs().max() + 1e-30))
    print(B, "on vs off:", {k.split(".conv")[0][-14:] + ("w" if "weight" in k else "b"): f"{err(b[k], a[k]):.1e}" for k in a})
    print(B, "off vs off:", max(err(c[k], a[k]) for k in a))
