#!/usr/bin/env python
"""Benchmark of the ConvLSTM recurrence hot path (contract: see DESIGN.md, "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], the largest single-GPU config the metric is quoted on):
  encoder-forecaster generator inference, 128x128 frames, hidden [64, 64], kernel 3, T = 10 -> 10,
  batch 32 sequences per GPU, bf16 tensor-core mode, synthetic radar-like frames, random-init weights.
One "step" = one batch through the front-end kernel -> 20 encoder cell steps -> 20 forecaster cell steps -> head
(42 kernel launches).  `--workload train` = configs[2]-style training step, `--workload radar` = configs[3].  N > 1: every rank runs its own batch (weak scaling, no data-path collective:
inference shards by batch).

Prints ONE JSON line (rank 0).  `value` = sequences/s with inputs resident in HBM; `e2e` = the same through
the public API with pinned-host frames in and predicted frames out inside the timed region; `roofline` =
the dominant kernel (fused tcgen05 cell step) against the measured bf16 peak; `cpu_baseline` = the oracle
port (same ATen CPU ops as the reference) on this box's host cores, bounded sample.
`--impl reference` times that CPU path alone (the reference is pure PyTorch and does not travel to the GPU
box, so the oracle port -- pinned to the reference by tests/golden -- stands in for it).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=32, H=128, W=128, hidden=[64, 64], k=3, t_in=10, t_out=10, in_channels=1)
WORKLOAD = ("cfg2: ConvLSTM encoder-forecaster generator inference, 128x128, hidden [64,64], k3, T=10->10, "
            "batch 32 per GPU")
RADAR_CFG = dict(B=16, H=256, W=256, hidden=[128, 128, 128], k=3, t_in=20, t_out=20, in_channels=1)
RADAR_WORKLOAD = ("cfg4 (BASELINE configs[3]): radar-scale nowcasting inference, 256x256, 3-layer ConvLSTM hidden "
                  "[128,128,128], k3, T=20->20, batch 16 per GPU")
FALLBACK_PEAK_TFLOPS = 1590.0   # B200_PROFILING.md fallback (burst)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "radar"],
                    help="infer = BASELINE configs[1] (default, the headline line); train = configs[2]-style recurrence "
                         "training step (global batch 64 sharded over the ranks, strong scaling); radar = configs[3] "
                         "(256x256, 3 layers of hidden 128, T=20->20, batch 16 per GPU), same line format as infer")
    ap.add_argument("--cpu-sample", type=int, default=1, help="sequences per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------- CPU baseline
def cpu_reference_run(steps: int, warmup: int, sample_b: int):
    """The reference's CPU path for this workload, via the oracle port (F.conv2d / sigmoid / tanh on oneDNN,
    fp32, all host threads).  Each step = `sample_b` sequences.  Returns (sequences/s, seconds/step, cores)."""
    import torch
    from oracle import convlstm_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    hd, k = CFG["hidden"], CFG["k"]

    def mk(o, i, kk):
        bound = 1.0 / (i * kk * kk) ** 0.5
        return (torch.rand(o, i, kk, kk) * 2 - 1) * bound, (torch.rand(o) * 2 - 1) * bound

    w_init, b_init = mk(hd[0], CFG["in_channels"] + 2, 3)
    dims = [hd[0]] + hd
    enc = [mk(4 * dims[l + 1], dims[l] + dims[l + 1], k) for l in range(len(hd))]
    fdims = [0] + hd
    fc = [mk(4 * fdims[l + 1], fdims[l] + fdims[l + 1], k) for l in range(len(hd))]
    w_head, b_head = mk(1, hd[-1], 1)
    frames = torch.rand(sample_b, CFG["t_in"], CFG["in_channels"], CFG["H"], CFG["W"])

    def one():
        with torch.no_grad():
            return O.nowcast_forward(frames, w_init, b_init, [w for w, _ in enc], [b for _, b in enc],
                                     [w for w, _ in fc], [b for _, b in fc], w_head, b_head, CFG["t_out"])

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = args.steps
    val, sec_per_step, cores = cpu_reference_run(steps, args.warmup, args.cpu_sample)
    sample = f"{args.cpu_sample} sequence(s) per step of the same workload (B reduced from {CFG['B']})"
    line = {
        "impl": "reference", "metric": "generator_inference_sequences_per_sec", "value": val, "unit": "sequences/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port of the reference path (torch CPU, oneDNN)"},
        "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------- our arm
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["bf16_tflops_sustained"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_PEAK_TFLOPS, FALLBACK_PEAK_TFLOPS, "fallback (B200_PROFILING.md)"


def committed_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("cell_fwd_cfg2_dram_bytes_per_launch")
        except Exception:
            return None
    return None


TRAIN = dict(global_batch=64, H=128, W=128, C=64, hidden=[64, 64], k=3, T=20)
TRAIN_WORKLOAD = ("cfg3-style: encoder-forecaster generator training step (front-end conv, 2-layer ConvLSTM hidden [64,64] "
                  "k3 encoder T=10 + forecaster T=10, 1x1 head, L1 frame loss): fwd + BPTT + grad all-reduce + clip + "
                  "Adam, 128x128, global batch 64 sharded by batch; no discriminator exists in the reference")


def run_train(args):
    """Training-step throughput of the recurrence (forward rollout, BPTT through plc_cell_bwd, bucketed gradient
    all-reduce overlapped with BPTT, clip 0.5, Adam) -- the reference order of trainer.py:290-315."""
    import torch
    import torch.distributed as dist
    import plconv
    from plconv.parallel import init_distributed, shard_batch
    from plconv.training import TrainStep

    rank, world, local = init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = TRAIN
    sl = shard_batch(cfg["global_batch"], rank, world)
    B = sl.stop - sl.start
    torch.manual_seed(1234)                       # identical weights on every rank
    gen = plconv.NowcastGenerator(1, cfg["hidden"], cfg["k"], cfg["T"] // 2, cfg["T"] // 2, "bf16").to(dev)
    torch.manual_seed(1234 + rank)
    frames = torch.relu(torch.randn(B, cfg["T"] // 2, 1, cfg["H"], cfg["W"], device=dev) + 0.3)
    tgt = torch.relu(torch.randn(B, cfg["T"] // 2, 1, cfg["H"], cfg["W"], device=dev) + 0.3)
    groups = [c.parameters() for c in gen.encoder.cells] + [c.parameters() for c in gen.forecaster.cells] + \
             [list(gen.init_conv.parameters()) + list(gen.head.parameters())]
    step = TrainStep(gen, groups, lr=5e-4, grad_clip_norm=0.5)

    def forward_loss():
        pred = gen(frames)                         # front-end conv -> encoder -> forecaster -> head, all in libplc.so
        return (pred - tgt).abs().mean()           # L1 on frames, as the reference's loss terms (combined_loss.py)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, Wm = args.steps, max(args.warmup, 3)
    for _ in range(Wm):
        step(forward_loss)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for _ in range(K):
        loss = step(forward_loss)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        L = len(cfg["hidden"])
        flops_fwd = gen.cell_flops_per_sequence(cfg["H"], cfg["W"])
        seqs = cfg["global_batch"] * K
        peak_sus, peak_burst, peak_src = measured_peaks()
        algo_tf = 3.0 * flops_fwd * seqs / (ms * 1e-3) / 1e12 / world
        print(json.dumps({
            "metric": "generator_train_sequences_per_sec", "value": seqs / (ms * 1e-3), "unit": "sequences/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": TRAIN_WORKLOAD, "global_batch": cfg["global_batch"], "per_gpu_batch": B,
                       "parallelism": f"dp{world} (batch shards + NCCL grad all-reduce overlapped with BPTT)",
                       "l2": "inputs larger than L2"},
            "gpu_launches": K * (cfg["T"] * L * (1 + 3) + 8),
            "loss": None if loss is None else float(loss),
            "roofline": {"bound": "tensor", "achieved": algo_tf, "peak": peak_sus, "unit": "TFLOP/s",
                         "frac": algo_tf / peak_sus, "peak_source": peak_src,
                         "note": "whole-step algorithmic 3*F_fwd per GPU (gate recompute not counted)", "traffic": None},
            "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def main():
    global CFG, WORKLOAD
    args = parse()
    if args.workload == "radar":
        CFG, WORKLOAD = RADAR_CFG, RADAR_WORKLOAD
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.workload == "train":
        run_train(args)
        return

    import torch
    import torch.distributed as dist
    import plconv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B, H, W = CFG["B"], CFG["H"], CFG["W"]
    torch.manual_seed(1234 + rank)
    model = plconv.NowcastGenerator(CFG["in_channels"], CFG["hidden"], CFG["k"], CFG["t_in"], CFG["t_out"],
                                    "bf16").to(dev)
    runner = plconv.NowcastRunner(model, B, H, W, dev)
    # radar-like non-negative frames (real rain is >= 0; fenhe_dataset.py:26-29,163-179)
    frames_host = torch.relu(torch.randn(B, CFG["t_in"], CFG["in_channels"], H, W) + 0.3).pin_memory()
    out_host = torch.empty(CFG["t_out"], B, H, W, dtype=torch.float32).pin_memory()
    frames_dev = frames_host.to(dev)
    K, Wm = args.steps, max(args.warmup, 3)

    # ---------------- device-resident region (value, roofline)
    for _ in range(Wm):
        runner.run(frames_dev)
    sampler = ClockSampler(local)
    events = []
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        runner.run(frames_dev, events=events)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * K / (elapsed_ms * 1e-3)

    # dominant kernel = fused cell step on the full shape (Cin = Ch = 64): flops / average launch duration
    full = [(a.elapsed_time(b), pw) for a, b, pw in events if pw.Cin == CFG["hidden"][0]]
    flops_full = 2.0 * B * H * W * (2 * CFG["hidden"][0]) * CFG["k"] ** 2 * 4 * CFG["hidden"][0]
    avg_ms = sum(t for t, _ in full) / len(full)
    achieved_tf = flops_full / (avg_ms * 1e-3) / 1e12
    peak_sus, peak_burst, peak_src = measured_peaks()
    cell_ms_total = sum(a.elapsed_time(b) for a, b, _ in events)

    # ---------------- end-to-end region: pinned host frames -> device -> rollout -> predicted frames -> host.
    # Every step copies ITS input from pinned host memory and ITS result back; the copies run on a side stream
    # (double-buffered) so that step i's D2H and step i+1's H2D overlap the compute of the neighbouring steps.
    copy_s = torch.cuda.Stream()
    main_s = torch.cuda.current_stream()
    in_buf = [torch.empty_like(frames_dev) for _ in range(2)]
    out_buf = [torch.empty_like(runner.out) for _ in range(2)]
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    run_done = [torch.cuda.Event() for _ in range(2)]
    d2h_done = [torch.cuda.Event() for _ in range(2)]

    def stage_in(i):
        s_ = i & 1
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(run_done[s_])            # the run that last read this input buffer has finished
            in_buf[s_].copy_(frames_host, non_blocking=True)
            h2d_done[s_].record(copy_s)

    def e2e_loop(n):
        stage_in(0)
        for i in range(n):
            s_ = i & 1
            if i + 1 < n:
                stage_in(i + 1)
            main_s.wait_event(h2d_done[s_])
            main_s.wait_event(d2h_done[s_])            # the D2H that last read this output buffer has finished
            out = runner.run(in_buf[s_], out=out_buf[s_])
            run_done[s_].record(main_s)
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(run_done[s_])
                out_host.copy_(out, non_blocking=True)
                d2h_done[s_].record(copy_s)
        main_s.wait_stream(copy_s)

    e2e_loop(Wm)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_loop(K)
    s1.record()
    torch.cuda.synchronize()
    barrier()
    e2e_ms = max_over_ranks(s0.elapsed_time(s1))
    e2e_value = world * B * K / (e2e_ms * 1e-3)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = cpu_reference_run(steps=3, warmup=1, sample_b=args.cpu_sample)
        cpu_baseline = {"value": v, "unit": "sequences/s", "cores": cores, "kind": "port",
                        "sample": f"3 timed passes of {args.cpu_sample} sequence(s) of the same workload "
                                  f"({sec:.2f} s per pass); oracle port = reference's ATen CPU ops"}

    if rank == 0:
        line = {
            "metric": "generator_inference_sequences_per_sec", "value": value, "unit": "sequences/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": elapsed_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": f"dp{world} (batch shards, no collective)",
                       "l2": f"inputs larger than L2 ({B * H * W * (2 * 2 + 4) * CFG['hidden'][0] / 1e6:.0f} MB of bf16 "
                             "x/h + fp32 c operands per cell step vs 126 MB L2)",
                       "cell_steps_per_sequence": runner.cell_launches_per_run},
            "e2e": {"value": e2e_value, "unit": "sequences/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": frames_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": K * runner.launches_per_run,
            "roofline": {"bound": "tensor", "kernel": f"conv_igemm_tc_kernel<256,EPI_LSTM_FWD> (fused cell step {CFG['hidden'][0]}->{CFG['hidden'][0]})",
                         "achieved": achieved_tf, "peak": peak_sus, "unit": "TFLOP/s", "frac": achieved_tf / peak_sus,
                         "frac_of_burst": achieved_tf / peak_burst, "peak_burst": peak_burst, "peak_source": peak_src,
                         "flops_per_launch": flops_full, "avg_launch_us": avg_ms * 1e3, "launches_timed": len(full),
                         "traffic": committed_traffic() if args.workload == "infer" else None,
                         "cell_kernels_share_of_step": cell_ms_total / elapsed_ms},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
