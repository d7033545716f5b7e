#!/usr/bin/env python
"""Pretty-print a bench.py JSON line (train workloads): headline, roofline, per-kind kernel breakdown."""
import json
import sys


def show(path):
    txt = open(path).read().strip().splitlines()
    lines = [l for l in txt if l.startswith("{")]
    if not lines:
        print(path, "NO JSON LINE")
        return
    d = json.loads(lines[-1])
    print(f"{path}: {d['metric']} n_gpus={d['n_gpus']} value={d['value']:.1f} ms/step={d['ms_per_step']:.2f} "
          f"e2e={d['e2e']['value']:.1f} launches/step={d.get('gpu_launches_per_step')} loss={d.get('loss')}")
    r = d.get("roofline") or {}
    if "avg_launch_us" in r:
        print(f"  roofline: {r['avg_launch_us']:.1f} us  {r['achieved']:.1f} TF  frac={r['frac']:.3f} burst={r['frac_of_burst']:.3f} "
              f"share={r.get('share_of_step', r.get('cell_kernels_share_of_step', 0)):.3f}")
    if r.get("cell_fwd"):
        f = r["cell_fwd"]
        print(f"  cell_fwd: {f['avg_launch_us']:.1f} us {f['achieved']:.1f} TF frac={f['frac']:.3f} burst={f['frac_of_burst']:.3f}")
    if r.get("step"):
        s = r["step"]
        print(f"  step: {s['algorithmic_tflops_per_gpu']:.1f} TF frac={s['frac']:.3f} burst={s['frac_of_burst']:.3f}")
    for k, v in (r.get("kernels") or {}).items():
        tf = v["executed_tflops"]
        print(f"    {k:14s} n={v['launches_per_step']:6.1f} avg={v['avg_launch_us']:9.1f} us share={v['share_of_step']:.3f} "
              f"TF={'-' if tf is None else round(tf, 1)}")
    print("  cpu:", d.get("cpu_baseline"))
    print("  clocks:", d.get("clocks"))


for p in sys.argv[1:]:
    show(p)
