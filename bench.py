#!/usr/bin/env python
"""Benchmark of the ConvLSTM recurrence hot path (contract: see DESIGN.md, "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|infer|radar|radar-train]

Default workload = BASELINE.json configs[2] ("cfg3"), the configuration the metric "GAN train sequences/sec at 1/2/4/8
B200" is quoted on: one TRAINING step of the encoder-forecaster ConvLSTM generator (front-end conv -> 2-layer ConvLSTM
encoder, T = 10 -> 2-layer forecaster, T = 10 -> 1x1 head; hidden [64, 64], k 3, 128x128 frames) against the
repo-defined discriminator (strided 2-D + 3-D convolutions on the same tensor-core core), bf16 tensor-core mode, global
batch 64 sharded by batch over the ranks (STRONG scaling), gradients all-reduced with NCCL and overlapped with BPTT.
One "step" = D step (real + detached fake clip) and G step (L1 + adversarial through D, BPTT through 40 cell steps),
in the reference's order zero_grad -> forward -> loss -> NaN-skip -> backward -> all-reduce -> clip 0.5 -> Adam
(src/training/trainer.py:290-315).  `--no-gan` drops the discriminator (generator-only L1 training).

Prints ONE JSON line (rank 0):
  value     sequences/s, inputs already resident in HBM when the timed region starts
  e2e       the same through the public API with PINNED HOST batches: every step's frames/targets are copied host ->
            device (DevicePrefetcher, side stream) and the step's loss is read back device -> host, inside the timed region
  roofline  the dominant call of the step -- the BPTT cell step (plc_cell_bwd = gate recompute + dgrad + wgrad) -- from
            live per-launch CUDA events (plc_timing_*, a separate pass so the events never sit in a reported region),
            algorithmic 2*F per call (the recomputed gate MMAs are NOT counted) against the
            measured bf16 peak; `kernels` holds every kernel kind's own time / executed-FLOP rate, `step` the whole step
  cpu_baseline  the oracle port of the same training step (same ATen CPU ops as the reference) on this box's host cores,
            bounded sample
`--impl reference` times that CPU path alone with the identical `config` (the reference is pure PyTorch and does not
travel to the GPU box, so the oracle port -- pinned to the reference by tests/golden -- stands in for it; its per-step
batch is a bounded sample of the workload, stated in `cpu_baseline.sample`; sequences/s is per-sequence work over time,
so it is the quantity that scales linearly in the batch on a CPU).
`--workload infer` = configs[1] (inference, batch 32/GPU, weak scaling), `--workload radar` = configs[3] inference,
`--workload radar-train` = configs[3] training (256x256, 3 x hidden 128, T = 20 -> 20, batch 16/GPU, weak scaling).
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

# cfg3 (BASELINE configs[2]): the configuration the headline metric is quoted on
TRAIN = dict(global_batch=64, H=128, W=128, hidden=[64, 64], k=3, t_in=10, t_out=10, in_channels=1, scaling="strong")
# cfg4 (BASELINE configs[3]) as a TRAINING step: 40-step BPTT through 3 x hidden 128 at 256x256 (state ring ~97 GB)
RADAR_TRAIN = dict(per_gpu_batch=16, H=256, W=256, hidden=[128, 128, 128], k=3, t_in=20, t_out=20, in_channels=1,
                   scaling="weak")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "infer", "radar", "radar-train"],
                    help="train = BASELINE configs[2] (default, the headline line: GAN training step, global batch 64, "
                         "strong scaling); infer = configs[1]; radar = configs[3] inference; radar-train = configs[3] "
                         "training (batch 16 per GPU, weak scaling)")
    ap.add_argument("--no-gan", action="store_true", help="train workloads: generator-only L1 training, no discriminator")
    ap.add_argument("--global-batch", type=int, default=None,
                    help="train: override the global batch (e.g. 8 on one GPU = the per-GPU shard of the 8-GPU run)")
    ap.add_argument("--no-graph", action="store_true",
                    help="train: issue every launch from Python instead of replaying the step as one CUDA graph")
    ap.add_argument("--cpu-sample", type=int, default=None, help="sequences per CPU-baseline step")
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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["bf16_tflops_sustained"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_PEAK_TFLOPS, FALLBACK_PEAK_TFLOPS, "fallback (B200_PROFILING.md)"


def committed_traffic(key: str):
    """dram bytes per launch of a kernel from the committed ncu --set full capture (profiles/roofline_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


def cfg3_traffic(which: str, saved: bool, per_gpu_batch: int, cfg):
    """dram bytes per launch / per BPTT call for the cfg3 cell shape (64 -> 64, 128x128, k3) from the committed ncu --set
    full captures at 64 sequences, scaled linearly to this run's per-GPU batch; None for other shapes (no capture)."""
    if cfg["H"] != 128 or cfg["W"] != 128 or list(cfg["hidden"]) != [64, 64]:
        return None
    v = committed_traffic(f"{which}_cfg3_b64_{'saved' if saved else 'recompute'}")
    return None if v is None else int(v * per_gpu_batch / 64)


# ================================================================================= training workloads (cfg3 / cfg4)
def train_cfg(args):
    cfg = RADAR_TRAIN if args.workload == "radar-train" else TRAIN
    if args.global_batch and cfg["scaling"] == "strong":
        cfg = dict(cfg, global_batch=args.global_batch)
    return cfg


def train_batch_sizes(cfg, world):
    """-> (global batch, per-GPU batch)"""
    if cfg["scaling"] == "strong":
        if cfg["global_batch"] % world:
            raise SystemExit(f"global batch {cfg['global_batch']} is not divisible by {world} GPUs")
        return cfg["global_batch"], cfg["global_batch"] // world
    return cfg["per_gpu_batch"] * world, cfg["per_gpu_batch"]


def train_config_dict(args, world):
    """The `config` object of the JSON line -- built the same way by both arms so the driver can compare them."""
    cfg = train_cfg(args)
    gb, pb = train_batch_sizes(cfg, world)
    gan = not args.no_gan
    name = "cfg4 (BASELINE configs[3])" if args.workload == "radar-train" else "cfg3 (BASELINE configs[2])"
    model = (f"encoder-forecaster ConvLSTM generator (coord front-end conv, {len(cfg['hidden'])}-layer ConvLSTM hidden "
             f"{cfg['hidden']} k{cfg['k']}, T={cfg['t_in']}->{cfg['t_out']}, 1x1 head)")
    if gan:
        model += (" + repo-defined clip discriminator (strided 2-D conv -> two strided 3-D convs -> 3x3 score conv, "
                  "LeakyReLU 0.2); losses: D = BCE(real,1) + BCE(fake,0), G = L1 + 0.05 * BCE(D(fake),1); neither the "
                  "encoder-forecaster nor a discriminator exists in the reference (SURVEY.md section 0)")
    step = ("D step + G step, each zero_grad -> forward -> loss -> NaN-skip -> backward (BPTT) -> grad all-reduce -> clip "
            "0.5 -> Adam (trainer.py:290-315)") if gan else \
           "zero_grad -> forward -> L1 loss -> NaN-skip -> backward (BPTT) -> grad all-reduce -> clip 0.5 -> Adam"
    return {"workload": f"{name}: {'GAN' if gan else 'generator'} training step, {cfg['H']}x{cfg['W']} frames, bf16 "
                        f"tensor-core mode; {model}; step = {step}",
            "global_batch": gb, "per_gpu_batch": pb,
            "parallelism": f"dp{world} (batch shards + NCCL gradient all-reduce overlapped with BPTT)",
            "l2": "inputs larger than L2 (every cell step streams >= 0.5 GB of state through a 126 MB L2)"}


def train_metric(args):
    return "generator_train_sequences_per_sec" if args.no_gan else "gan_train_sequences_per_sec"


def cpu_train_run(args, steps: int, warmup: int, sample_b: int):
    """The reference's CPU path for the training step, via the oracle port (F.conv2d / sigmoid / tanh on oneDNN, fp32
    autograd, all host threads): forward, loss, backward, clip 0.5, Adam -- on `sample_b` sequences per step.
    Returns (sequences/s, seconds/step, cores)."""
    import torch
    from oracle import convlstm_oracle as O
    cfg = train_cfg(args)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    hd, k = cfg["hidden"], cfg["k"]

    def mk(o, i, kk):
        bound = 1.0 / (i * kk * kk) ** 0.5
        w = ((torch.rand(o, i, kk, kk) * 2 - 1) * bound).requires_grad_()
        b = ((torch.rand(o) * 2 - 1) * bound).requires_grad_()
        return w, b

    w_init, b_init = mk(hd[0], cfg["in_channels"] + 2, 3)
    dims = [hd[0]] + hd
    enc = [mk(4 * dims[l + 1], dims[l] + dims[l + 1], k) for l in range(len(hd))]
    fdims = [0] + hd
    fc = [mk(4 * fdims[l + 1], fdims[l] + fdims[l + 1], k) for l in range(len(hd))]
    w_head, b_head = mk(1, hd[-1], 1)
    g_params = [w_init, b_init, w_head, b_head] + [t for wb in enc + fc for t in wb]
    g_opt = torch.optim.Adam(g_params, lr=5e-4)
    frames = torch.relu(torch.randn(sample_b, cfg["t_in"], cfg["in_channels"], cfg["H"], cfg["W"]) + 0.3)
    target = torch.relu(torch.randn(sample_b, cfg["t_out"], 1, cfg["H"], cfg["W"]) + 0.3)
    disc = None
    if not args.no_gan:
        from oracle import gan_oracle as G
        disc = G.make_discriminator_params(seed=4321)
        for t in disc.values():
            t.requires_grad_()
        d_opt = torch.optim.Adam(list(disc.values()), lr=2e-4, betas=(0.5, 0.999))

    def one():
        fake = O.nowcast_forward(frames, w_init, b_init, [w for w, _ in enc], [b for _, b in enc],
                                 [w for w, _ in fc], [b for _, b in fc], w_head, b_head, cfg["t_out"])
        if disc is not None:
            d_opt.zero_grad()
            clips = torch.cat([torch.cat([frames, target], 1), torch.cat([frames, fake.detach()], 1)], 0)
            d_loss = G.d_loss(G.discriminator_forward(clips, disc), sample_b)
            d_loss.backward()
            torch.nn.utils.clip_grad_norm_(list(disc.values()), 0.5)
            d_opt.step()
        g_opt.zero_grad()
        loss = (fake - target).abs().mean()
        if disc is not None:
            for t in disc.values():
                t.requires_grad_(False)
            loss = loss + 0.05 * G.g_adv_loss(G.discriminator_forward(torch.cat([frames, fake], 1), disc))
        loss.backward()
        if disc is not None:
            for t in disc.values():
                t.requires_grad_(True)
        torch.nn.utils.clip_grad_norm_(g_params, 0.5)
        g_opt.step()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps, cores


def run_train_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cfg = train_cfg(args)
    sample_b = args.cpu_sample or (1 if args.workload == "radar-train" else 2)
    val, sec, cores = cpu_train_run(args, args.steps, args.warmup, sample_b)
    sample = (f"{sample_b} sequence(s) per step of the same training step ({sec:.2f} s per step): forward + loss + "
              f"autograd BPTT + clip + Adam through the oracle port (the reference's ATen CPU ops, fp32)")
    print(json.dumps({
        "impl": "reference", "metric": train_metric(args), "value": val, "unit": "sequences/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config_dict(args, world),
        "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_train(args):
    """Training-step throughput through the public API (plconv.NowcastGenerator / Discriminator / GanTrainStep)."""
    import torch
    import torch.distributed as dist
    import plconv
    from plconv import _lib
    from plconv.parallel import init_distributed
    from plconv.trainer import DevicePrefetcher

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    rank, world, local = init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = train_cfg(args)
    gb, B = train_batch_sizes(cfg, world)
    H, W, t_in, t_out = cfg["H"], cfg["W"], cfg["t_in"], cfg["t_out"]
    gan = not args.no_gan

    torch.manual_seed(1234)                       # identical initial weights on every rank
    gen = plconv.NowcastGenerator(cfg["in_channels"], cfg["hidden"], cfg["k"], t_in, t_out, "bf16").to(dev)
    if gan:
        from plconv.gan import Discriminator, GanTrainStep
        disc = Discriminator().to(dev)
        step = GanTrainStep(gen, disc, lr_g=5e-4, lr_d=2e-4, lambda_adv=0.05, grad_clip_norm=0.5)
    else:
        from plconv.training import TrainStep
        groups = [list(c.parameters()) for c in gen.forecaster.cells] + [list(c.parameters()) for c in gen.encoder.cells] + \
                 [list(gen.init_conv.parameters()) + list(gen.head.parameters())]
        ts = TrainStep(gen, groups, lr=5e-4, grad_clip_norm=0.5)

        def step(frames, target):
            return ts(lambda: (gen(frames) - target).abs().mean())     # L1 on frames (combined_loss.py's |.| terms)

    # synthetic radar-like non-negative frames (real rain is >= 0; fenhe_dataset.py:26-29,163-179), pinned host memory
    gcpu = torch.Generator().manual_seed(1234 + rank)
    n_host = 3
    host = [(torch.relu(torch.randn(B, t_in, cfg["in_channels"], H, W, generator=gcpu) + 0.3).pin_memory(),
             torch.relu(torch.randn(B, t_out, 1, H, W, generator=gcpu) + 0.3).pin_memory()) for _ in range(n_host)]
    frames_dev, target_dev = (t.to(dev) for t in host[0])

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

    K, Wm = args.steps, max(args.warmup, 3)
    eager_step = step
    if not args.no_graph:
        # the whole step (D step + G step: forward, BPTT, NCCL all-reduce, clip, Adam) recorded once, replayed per batch
        from plconv.training import GraphedStep
        step = GraphedStep(eager_step, (frames_dev, target_dev), warmup=3)
    # ---------------- device-resident region (value)
    for _ in range(Wm):
        step(frames_dev, target_dev)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for _ in range(K):
        loss = step(frames_dev, target_dev)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = gb * K / (ms * 1e-3)
    from plconv import nn as _pnn
    saved_bytes = _pnn.LAST_SAVED_GATES_BYTES        # what the measured step did (decided when it was recorded)

    # ---------------- end-to-end region: pinned host batches -> DevicePrefetcher (side stream, one batch ahead) ->
    # step -> the step's loss back to pinned host memory; every step copies ITS batch in and ITS result out.
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def e2e_loop(n):
        for fr, tg in DevicePrefetcher((host[i % n_host] for i in range(n)), dev):
            out = step(fr, tg)
            loss_host.copy_(out.detach().reshape(1).float(), non_blocking=True)

    e2e_loop(Wm)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_loop(K)
    s1.record()
    torch.cuda.synchronize()
    barrier()
    e2e_ms = max_over_ranks(s0.elapsed_time(s1))
    e2e_value = gb * K / (e2e_ms * 1e-3)
    last_loss = float(loss_host.item())

    # ---------------- per-launch pass (separate from the reported regions): CUDA events around every kernel the library
    # launches, tagged with its kind and algorithmic FLOPs (plc_timing_*)
    n_prof = 2
    torch.cuda.synchronize()
    if hasattr(step, "close"):
        step.close()                                # the graph's private pool (rings, saved gates) goes back to the driver:
                                                    # at radar scale the eager pass below needs that memory again
    _lib.timing_enable(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(n_prof):
        eager_step(frames_dev, target_dev)         # eager: same kernels as the graph, each bracketed by an event pair
    p1.record()
    torch.cuda.synchronize()
    rec = _lib.timing_collect()
    _lib.timing_enable(False)
    prof_ms = p0.elapsed_time(p1)
    barrier()

    if rank == 0:
        peak_sus, peak_burst, peak_src = measured_peaks()
        kinds = {}
        for kind, t_ms, fl in rec:
            d = kinds.setdefault(kind, {"launches": 0, "ms": 0.0, "flops": 0.0})
            d["launches"] += 1
            d["ms"] += t_ms
            d["flops"] += fl
        kernels = {}
        for kind, d in sorted(kinds.items(), key=lambda kv: -kv[1]["ms"]):
            kernels[kind] = {"launches_per_step": d["launches"] / n_prof, "avg_launch_us": d["ms"] / d["launches"] * 1e3,
                             "share_of_step": (d["ms"] / n_prof) / (ms / K),
                             "executed_tflops": (d["flops"] / (d["ms"] * 1e-3) / 1e12) if d["flops"] else None}
        launches_per_step = len(rec) / n_prof
        # dominant call: the BPTT cell step = the three kernels of one plc_cell_bwd (gate recompute, dgrad, wgrad);
        # algorithmic FLOPs = dgrad + wgrad (the recomputed gate contraction is not counted)
        bwd = [kinds.get(k_) for k_ in ("bwd_gates", "bwd_dgrad", "bwd_wgrad")]
        roofline = None
        if all(bwd):
            calls = bwd[0]["launches"]
            bwd_ms = sum(d["ms"] for d in bwd)
            algo = bwd[1]["flops"] + bwd[2]["flops"]
            ach = algo / (bwd_ms * 1e-3) / 1e12
            fwd = kinds.get("cell_fwd")
            cell_flops_step = gen.cell_flops_per_sequence(H, W) * B          # forward F of every cell step, this GPU
            step_tf = 3.0 * cell_flops_step / (ms / K * 1e-3) / 1e12
            roofline = {
                "bound": "tensor",
                "kernel": "BPTT cell step = plc_cell_bwd[_saved]: conv_igemm_tc_kernel<256,EPI_LSTM_BWD_GATES> (gate "
                          "gradients dZ: from saved gates when `bptt_gates` says so, else gate recompute) + "
                          "conv_igemm_tc_kernel<128,EPI_PLAIN> (dgrad) + wgrad_tc_kernel2 (dW, db)",
                "achieved": ach, "peak": peak_sus, "unit": "TFLOP/s", "frac": ach / peak_sus,
                "frac_of_burst": ach / peak_burst, "peak_burst": peak_burst, "peak_source": peak_src,
                "flops_per_launch": algo / calls, "avg_launch_us": bwd_ms / calls * 1e3, "launches_timed": calls,
                "share_of_step": (bwd_ms / n_prof) / (ms / K),
                "note": "algorithmic FLOPs = dgrad + wgrad of the call (2F at Cin = Ch); the gate-recompute MMAs it also "
                        "executes are not counted.  Per-launch CUDA events from a separate eager pass of "
                        f"{n_prof} steps (never inside a reported region); share_of_step = the kind's event time per "
                        "step / the reported (graph-replayed) step time; peak = sustained bf16 matmul (the kernels run "
                        "inside a multi-second step), burst alongside",
                "traffic": cfg3_traffic("cell_bwd", saved_bytes > 0, B, cfg),
                "cell_fwd": None if not fwd else {
                    "kernel": "conv_igemm_tc_kernel<256,EPI_LSTM_FWD,2> (fused cell step, full K loop; zero-state first "
                              "steps are timed separately as cell_fwd_zero)",
                    "avg_launch_us": fwd["ms"] / fwd["launches"] * 1e3,
                    "achieved": fwd["flops"] / (fwd["ms"] * 1e-3) / 1e12,
                    "frac": fwd["flops"] / (fwd["ms"] * 1e-3) / 1e12 / peak_sus,
                    "frac_of_burst": fwd["flops"] / (fwd["ms"] * 1e-3) / 1e12 / peak_burst,
                    "launches_timed": fwd["launches"],
                    "traffic": cfg3_traffic("cell_fwd", saved_bytes > 0, B, cfg)},
                "step": {"algorithmic_tflops_per_gpu": step_tf, "frac": step_tf / peak_sus,
                         "frac_of_burst": step_tf / peak_burst,
                         "note": "3 * F_fwd of every cell step of the rollout / device-resident step time (cells only: "
                                 "front-end, head, discriminator, optimizer time is in the denominator, their FLOPs "
                                 "are not in the numerator)"},
                "kernels": kernels}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_b = args.cpu_sample or (1 if args.workload == "radar-train" else 2)
        v, sec, cores = cpu_train_run(args, steps=3, warmup=1, sample_b=sample_b)
        cpu_baseline = {"value": v, "unit": "sequences/s", "cores": cores, "kind": "port",
                        "sample": f"3 timed steps of {sample_b} sequence(s) of the same training step ({sec:.2f} s per "
                                  f"step); oracle port = the reference's ATen CPU ops, fp32 autograd"}

    if rank == 0:
        h2d = sum(t.numel() * t.element_size() for t in host[0])
        print(json.dumps({
            "metric": train_metric(args), "value": value, "unit": "sequences/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": train_config_dict(args, world),
            "e2e": {"value": e2e_value, "unit": "sequences/s", "ms_per_step": e2e_ms / K, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "bptt_gates": (f"saved by the forward pass ({saved_bytes / 2**30:.1f} GiB per rollout): the backward gate kernel "
                           "streams them instead of re-running the gate contraction" if saved_bytes else
                           "recomputed in the backward gate kernel (no extra memory)"),
            "launch_mode": "eager (one Python -> C ABI call per kernel)" if args.no_graph else
                           "the whole step replayed as one CUDA graph (plconv.training.GraphedStep)",
            "gpu_launches": int(round(K * launches_per_step)),
            "gpu_launches_per_step": launches_per_step,
            "loss": last_loss,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks}), flush=True)
    shutdown(step, world, dev)


def shutdown(step, world, dev):
    """Orderly exit of a multi-rank run: the JSON line is already out.  Release the CUDA graph (it holds captured NCCL
    kernels) before the process group goes away, and never let a stuck communicator teardown keep the job alive."""
    import torch
    import torch.distributed as dist
    if world <= 1:
        return
    watchdog = threading.Timer(45.0, lambda: os._exit(0))     # the measurement is complete; do not hang the launcher
    watchdog.daemon = True
    watchdog.start()
    torch.cuda.synchronize(dev)
    dist.barrier()
    if hasattr(step, "close"):
        step.close()
    torch.cuda.synchronize(dev)
    dist.destroy_process_group()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)      # skip interpreter teardown (NCCL watchdog threads, graph pools): everything is reported and released


# ================================================================================= inference workloads (cfg2 / cfg4)
def cpu_reference_run(steps: int, warmup: int, sample_b: int):
    """The reference's CPU path for the inference workload, via the oracle port (F.conv2d / sigmoid / tanh on oneDNN,
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


def infer_config_dict(world):
    B, H, W = CFG["B"], CFG["H"], CFG["W"]
    return {"workload": WORKLOAD, "global_batch": world * B, "parallelism": f"dp{world} (batch shards, no collective)",
            "l2": f"inputs larger than L2 ({B * H * W * (2 * 2 + 4) * CFG['hidden'][0] / 1e6:.0f} MB of bf16 "
                  "x/h + fp32 c operands per cell step vs 126 MB L2)",
            "cell_steps_per_sequence": len(CFG["hidden"]) * (CFG["t_in"] + CFG["t_out"])}


def run_infer_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    sample_b = args.cpu_sample or 1
    val, sec_per_step, cores = cpu_reference_run(args.steps, args.warmup, sample_b)
    sample = f"{sample_b} sequence(s) per step of the same workload (B reduced from {CFG['B']})"
    print(json.dumps({
        "impl": "reference", "metric": "generator_inference_sequences_per_sec", "value": val, "unit": "sequences/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": infer_config_dict(world),
        "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_infer(args):
    import torch
    import torch.distributed as dist
    import plconv
    from plconv import _lib

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

    # ---------------- device-resident region (value): no per-launch events in here
    for _ in range(Wm):
        runner.run(frames_dev)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        runner.run(frames_dev)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * K / (elapsed_ms * 1e-3)

    # ---------------- per-launch pass (separate): the dominant kernel = fused cell step with the FULL K loop (x and h
    # taps); the zero-state first steps of the encoder run half of it and are recorded under their own kind
    n_prof = max(2, min(K, 5))
    _lib.timing_enable(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(n_prof):
        runner.run(frames_dev)
    p1.record()
    torch.cuda.synchronize()
    rec = _lib.timing_collect()
    _lib.timing_enable(False)
    prof_ms = p0.elapsed_time(p1)
    f_full = 2.0 * B * H * W * (2 * CFG["hidden"][-1]) * CFG["k"] ** 2 * 4 * CFG["hidden"][-1]
    full = [(t, fl) for kind, t, fl in rec if kind == "cell_fwd" and abs(fl - f_full) < 1e-3 * f_full]
    avg_ms = sum(t for t, _ in full) / len(full)
    achieved_tf = f_full / (avg_ms * 1e-3) / 1e12
    cell_ms_total = sum(t for kind, t, _ in rec if kind.startswith("cell_fwd"))
    peak_sus, peak_burst, peak_src = measured_peaks()

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
        sb = args.cpu_sample or 1
        v, sec, cores = cpu_reference_run(steps=3, warmup=1, sample_b=sb)
        cpu_baseline = {"value": v, "unit": "sequences/s", "cores": cores, "kind": "port",
                        "sample": f"3 timed passes of {sb} sequence(s) of the same workload "
                                  f"({sec:.2f} s per pass); oracle port = reference's ATen CPU ops"}

    if rank == 0:
        hd = CFG["hidden"][-1]
        print(json.dumps({
            "metric": "generator_inference_sequences_per_sec", "value": value, "unit": "sequences/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": elapsed_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": infer_config_dict(world),
            "e2e": {"value": e2e_value, "unit": "sequences/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": frames_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": K * runner.launches_per_run,
            "roofline": {"bound": "tensor", "kernel": f"conv_igemm_tc_kernel<256,EPI_LSTM_FWD> (fused cell step {hd}->{hd}, "
                                                      "full K loop; zero-state launches excluded)",
                         "achieved": achieved_tf, "peak": peak_sus, "unit": "TFLOP/s", "frac": achieved_tf / peak_sus,
                         "frac_of_burst": achieved_tf / peak_burst, "peak_burst": peak_burst, "peak_source": peak_src,
                         "flops_per_launch": f_full, "avg_launch_us": avg_ms * 1e3, "launches_timed": len(full),
                         "traffic": committed_traffic("cell_fwd_cfg2_dram_bytes_per_launch")
                         if args.workload == "infer" else None,
                         "cell_kernels_share_of_step": cell_ms_total / prof_ms},
            "cpu_baseline": cpu_baseline, "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def main():
    global CFG, WORKLOAD
    args = parse()
    if args.workload == "radar":
        CFG, WORKLOAD = RADAR_CFG, RADAR_WORKLOAD
    train = args.workload in ("train", "radar-train")
    if args.impl == "reference":
        (run_train_reference if train else run_infer_reference)(args)
    else:
        (run_train if train else run_infer)(args)


if __name__ == "__main__":
    main()
