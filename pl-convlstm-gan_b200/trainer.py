"""Training-loop wrapper around the drop-in Generator (SURVEY.md section 8f "next-3" / "next-4").

Mirrors ``src/training/trainer.py``: model / Adam / ReduceLROnPlateau / CombinedLoss set-up (:140-178), the
``train_epoch`` body (:286-329: zero_grad -> forward -> loss -> NaN-skip -> backward -> clip 0.5 -> Adam step ->
station RMSE), ``validate`` (:182-223), the epoch loop with scheduler, early stopping and the ``best_model.pth``
checkpoint dictionary (:338-418) -- and adds what a B200 job needs and the reference lacks: one process per GPU with
overlapped gradient all-reduce, pinned-memory double-buffered H2D, NO per-step host synchronisation (loss terms and
RMSE are accumulated on the device and read once per epoch; the NaN-skip is a device flag consumed by fused Adam),
and a resume path.

Out of scope: YAML config, GIS dataset, plots.  ``Trainer`` takes any iterable of
``(lr, dem, lu, s_coords, s_values)`` batches, the tuple ``FenheDataset`` yields (fenhe_dataset.py).
"""
from __future__ import annotations

import os
from dataclasses import asdict, dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .parallel import GradReducer

TERMS = ("total", "point", "conserve", "smooth", "temporal")


@dataclass
class TrainerConfig:
    """The fields of configs/default.yaml the loop reads (config.model.* / config.training.*)."""
    hidden_dims: Sequence[int] = (16, 32)
    lu_channels: int = 0
    scale_factor: int = 8
    learning_rate: float = 5e-4                 # trainer.py:155-158
    scheduler_factor: float = 0.7               # trainer.py:160-165
    scheduler_patience: int = 10
    grad_clip_norm: float = 0.5                 # trainer.py:311-314
    lambda_point: float = 1.0
    lambda_conserve: float = 1.0
    lambda_smooth: float = 0.1
    lambda_temporal: float = 0.05
    use_weighted_loss: bool = True
    weight_strategy: str = "log"
    epochs: int = 1
    early_stopping_patience: Optional[int] = None
    early_stopping_min_delta: float = 0.0
    mode: str = "bf16"                          # "bf16" tensor-core path | "fp32" validation mode
    # The reference builds Adam BEFORE the first forward creates ``upsample_blocks`` (generator.py:129-130), so those
    # parameters are never updated nor zeroed, but still clipped.  False reproduces that exactly;
    # True (default) trains them.
    optimizer_sees_upsample: bool = True
    output_dir: Optional[str] = None


class EarlyStopping:
    """src/utils/early_stopping.py:9-75: ``__call__(score, epoch) -> is_best``; sets ``early_stop`` after ``patience``
    epochs without an improvement larger than ``min_delta``."""

    def __init__(self, patience: int = 20, min_delta: float = 0.0, mode: str = "min"):
        self.patience, self.min_delta, self.mode = patience, min_delta, mode
        self.counter, self.best_score, self.early_stop, self.best_epoch = 0, None, False, 0

    def _better(self, new, best):
        return new < best - self.min_delta if self.mode == "min" else new > best + self.min_delta

    def __call__(self, score: float, epoch: int) -> bool:
        if self.best_score is None or self._better(score, self.best_score):
            self.best_score, self.best_epoch, self.counter = score, epoch, 0
            return True
        self.counter += 1
        if self.counter >= self.patience:
            self.early_stop = True
        return False

    def state_dict(self):
        return {"counter": self.counter, "best_score": self.best_score, "early_stop": self.early_stop,
                "best_epoch": self.best_epoch}

    def load_state_dict(self, sd):
        self.counter, self.best_score = sd["counter"], sd["best_score"]
        self.early_stop, self.best_epoch = sd["early_stop"], sd["best_epoch"]


def station_rmse(fake_hr: torch.Tensor, s_coords: torch.Tensor, s_values: torch.Tensor, scale_factor: float):
    """trainer.py:225-268 without boolean-mask indexing (no host sync): RMSE over valid, on-grid gauges; 0 if none."""
    B, T, _, H, W = fake_hr.shape
    coords = s_coords[0] if s_coords.dim() == 3 else s_coords
    sc = ((coords.float() + 0.5) * scale_factor - 0.5).long()
    rows, cols = sc[:, 0], sc[:, 1]
    on_grid = (rows >= 0) & (rows < H) & (cols >= 0) & (cols < W)
    at = fake_hr[:, :, 0][:, :, rows.clamp(0, H - 1), cols.clamp(0, W - 1)].float()
    obs = s_values if s_values.dim() == 3 else s_values.unsqueeze(0).expand(B, -1, -1)
    m = on_grid.view(1, 1, -1) & ~torch.isnan(obs)
    se = torch.where(m, (at - torch.nan_to_num(obs)) ** 2, torch.zeros_like(at)).sum()
    n = m.sum()
    return torch.where(n > 0, torch.sqrt(se / n.clamp(min=1)), torch.zeros_like(se))


class DevicePrefetcher:
    """Pinned host batches -> two fixed sets of device buffers, filled on a side stream one batch ahead of the compute
    stream.  The buffers are allocated once (per shape), so a step never touches the caching allocator across streams;
    a slot is refilled only after the compute stream has finished the step that read it."""

    def __init__(self, loader: Iterable, device: torch.device):
        self.loader, self.device = loader, device
        self.stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self.slots = [None, None]
        self.done = [None, None]

    def _stage(self, batch, s: int):
        if self.stream is None:
            return tuple(t.to(self.device) for t in batch), None, s
        bufs = self.slots[s]
        if bufs is None or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, batch)):
            bufs = self.slots[s] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch)
        with torch.cuda.stream(self.stream):
            if self.done[s] is not None:
                self.stream.wait_event(self.done[s])
            else:
                self.stream.wait_stream(torch.cuda.current_stream(self.device))   # the allocation above
            for b, t in zip(bufs, batch):
                b.copy_(t if t.is_cuda or t.is_pinned() else t.pin_memory(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return bufs, ev, s

    def __iter__(self):
        it = iter(self.loader)
        first = next(it, None)
        staged = self._stage(first, 0) if first is not None else None
        while staged is not None:
            cur, ev, s = staged
            nxt = next(it, None)
            staged = self._stage(nxt, 1 - s) if nxt is not None else None
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
            yield cur
            if ev is not None:
                self.done[s] = torch.cuda.Event()
                self.done[s].record(torch.cuda.current_stream(self.device))


class Trainer:
    def __init__(self, config: TrainerConfig, device=None, model: Optional[torch.nn.Module] = None,
                 loss_module: Optional[torch.nn.Module] = None, process_group=None):
        self.config = config
        self.device = torch.device(device if device is not None else "cuda")
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.history: Dict[str, List[float]] = {k: [] for k in ("epoch", "total_loss", "point_loss", "conserve_loss",
                                                                "smooth_loss", "temporal_loss", "rmse",
                                                                "learning_rate")}          # trainer.py:48-57
        self.best_rmse, self.best_epoch, self.start_epoch, self.skipped = float("inf"), 0, 0, 0
        self.early_stopping = (EarlyStopping(config.early_stopping_patience, config.early_stopping_min_delta)
                               if config.early_stopping_patience else None)
        self._setup_model(model, loss_module)

    # ---- trainer.py:140-178 ------------------------------------------------------------------------------------
    def _setup_model(self, model, loss_module):
        cfg = self.config
        if model is None:
            from .generator import Generator
            model = Generator(in_channels=1, dem_channels=1, lu_channels=cfg.lu_channels,
                              hidden_dims=list(cfg.hidden_dims), scale_factor=cfg.scale_factor,
                              mode=cfg.mode).to(self.device)
        if loss_module is None:
            from .losses import CombinedLoss
            loss_module = CombinedLoss(cfg.lambda_point, cfg.lambda_conserve, cfg.lambda_smooth, cfg.lambda_temporal,
                                       cfg.use_weighted_loss, cfg.weight_strategy)
        self.model, self.loss_module = model, loss_module
        before = list(model.parameters())                        # what the reference's optimizer sees
        if hasattr(model, "materialize"):
            model.materialize(cfg.scale_factor, self.device)
        every = list(model.parameters())
        seen = {id(p) for p in before}
        self.frozen = [] if cfg.optimizer_sees_upsample else [p for p in every if id(p) not in seen]
        self.trainable = every if cfg.optimizer_sees_upsample else before
        self._sync_initial_parameters()
        # one gradient bucket per top-level submodule: its all-reduce starts when its last gradient is ready
        groups, claimed = [], {id(p) for p in self.frozen}
        for child in model.children():
            grp = [p for p in child.parameters() if id(p) not in claimed and p.requires_grad]
            claimed.update(id(p) for p in grp)
            if grp:
                groups.append(grp)
        rest = [p for p in self.trainable if id(p) not in claimed and p.requires_grad]
        if rest:
            groups.append(rest)
        self.reducer = GradReducer(groups, self.pg)
        # quirk mode: gradients of the never-optimised parameters accumulate from step to step (nothing zeroes them)
        # and take part in the clip norm; under data parallelism they are averaged like the others so that every rank
        # computes the same clip coefficient
        self.frozen_reducer = GradReducer([self.frozen], self.pg) if self.frozen else None
        on_cuda = all(p.is_cuda for p in self.trainable)
        self.optimizer = torch.optim.Adam(self.trainable, lr=cfg.learning_rate, fused=on_cuda)
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(
            self.optimizer, mode="min", factor=cfg.scheduler_factor, patience=cfg.scheduler_patience)
        # fused Adam honours optimizer.found_inf on the device.  Quirk mode keeps the host check: a skipped batch must
        # not reach backward, or its NaN gradients would stay in the never-zeroed buffers for good
        self._device_skip = on_cuda and not self.frozen

    def _sync_initial_parameters(self):
        if self.world > 1:
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=0, group=self.pg)

    # ---- one optimisation step: trainer.py:290-315 ---------------------------------------------------------------
    def train_step(self, batch):
        lr, dem, lu, s_coords, s_values = batch
        self.reducer.zero_grad()                                                 # optimizer.zero_grad()
        fake_hr = self.model(lr, dem, lu)
        scale_factor = fake_hr.shape[-2] / lr.shape[-2]                          # trainer.py:299-301
        loss, parts = self.loss_module(fake_hr, lr, s_coords, s_values, scale_factor)
        bad = (~torch.isfinite(loss.detach())).to(torch.float32).reshape(1)      # trainer.py:306 `isnan -> continue`
        if self.world > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.pg)            # every rank takes the same branch
        if not self._device_skip and bool(bad.item()):
            self.skipped += 1
            return None
        loss.backward()
        self.reducer.finish()
        if self.frozen_reducer is not None:
            self._reduce_frozen()
        torch.nn.utils.clip_grad_norm_(self.trainable + self.frozen, self.config.grad_clip_norm)
        if self.frozen_reducer is not None and self.world > 1:
            self._frozen_prev = self.frozen_reducer.buckets[0]["flat"].clone()   # after the clip rescaled it
        if self._device_skip:            # no host sync: fused Adam skips the update when found_inf != 0 (as GradScaler does)
            self.optimizer.grad_scale, self.optimizer.found_inf = None, bad.reshape(())
            self.optimizer.step()
            del self.optimizer.grad_scale, self.optimizer.found_inf
        else:
            self.optimizer.step()
        with torch.no_grad():
            rmse = station_rmse(fake_hr.detach(), s_coords, s_values, scale_factor)
            ok = 1.0 - bad
            row = torch.stack([loss.detach().float(), parts["point"].float(), parts["conserve"].float(),
                               parts["smooth"].float(), parts["temporal"].float(), rmse.float()])
            return torch.cat([torch.nan_to_num(row) * ok, ok])                   # [total, 4 terms, rmse, counted]

    def _reduce_frozen(self):
        """Quirk mode under data parallelism: average only THIS step's contribution of the never-zeroed gradients."""
        b = self.frozen_reducer.buckets[0]
        if self.world > 1:
            prev = getattr(self, "_frozen_prev", None)
            if prev is None:
                prev = torch.zeros_like(b["flat"])
            if b["handle"] is not None:
                b["handle"].wait()                  # the hook reduced (prev + g_r) summed over ranks
                b["flat"].sub_(prev * self.world).div_(self.world).add_(prev)
        b["pending"], b["handle"] = len(b["params"]), None

    # ---- trainer.py:277-336 --------------------------------------------------------------------------------------
    def train_epoch(self, loader: Iterable) -> Dict[str, float]:
        self.model.train()
        acc = torch.zeros(7, dtype=torch.float32, device=self.device)
        steps = 0
        for batch in DevicePrefetcher(loader, self.device):
            row = self.train_step(batch)
            steps += 1
            if row is not None:
                acc += row
        if self.world > 1:
            dist.all_reduce(acc, group=self.pg)
        vals = acc.tolist()                                                       # the epoch's only host read
        n = max(vals[6], 1.0)
        if self._device_skip:
            self.skipped += steps * self.world - int(round(vals[6]))
        out = {k: vals[i] / n for i, k in enumerate(TERMS)}                       # np.mean per term, trainer.py:334
        out["rmse"] = vals[5] / n
        return out

    # ---- trainer.py:182-223 --------------------------------------------------------------------------------------
    def validate(self, loader: Optional[Iterable]) -> Optional[Dict[str, float]]:
        if loader is None:
            return None
        self.model.eval()
        acc = torch.zeros(3, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for lr, dem, lu, s_coords, s_values in DevicePrefetcher(loader, self.device):
                fake_hr = self.model(lr, dem, lu)
                sf = fake_hr.shape[-2] / lr.shape[-2]
                loss, _ = self.loss_module(fake_hr, lr, s_coords, s_values, sf)
                acc += torch.stack([loss.float(), station_rmse(fake_hr, s_coords, s_values, sf).float(),
                                    torch.ones((), device=self.device)])
        if self.world > 1:
            dist.all_reduce(acc, group=self.pg)
        self.model.train()
        loss, rmse, n = acc.tolist()
        return {"loss": loss / max(n, 1.0), "rmse": rmse / max(n, 1.0)}

    # ---- trainer.py:338-418 --------------------------------------------------------------------------------------
    def fit(self, train_loader: Iterable, val_loader: Optional[Iterable] = None) -> Dict[str, List[float]]:
        for epoch in range(self.start_epoch, self.config.epochs):
            avg = self.train_epoch(train_loader)
            self.history["epoch"].append(epoch)
            for k in TERMS:
                self.history[k + "_loss"].append(avg[k])
            self.history["rmse"].append(avg["rmse"])
            self.history["learning_rate"].append(self.optimizer.param_groups[0]["lr"])
            val = self.validate(val_loader)
            current = val["rmse"] if val else avg["rmse"]                        # trainer.py:362-368, 378-381
            self.scheduler.step(current)
            if self.early_stopping is not None:
                is_best = self.early_stopping(current, epoch)
            else:
                is_best = current < self.best_rmse
            self.start_epoch = epoch + 1
            if is_best:
                self.best_rmse, self.best_epoch = current, epoch
                if self.config.output_dir and self.rank == 0:
                    self.save_checkpoint(os.path.join(self.config.output_dir, "best_model.pth"), epoch, current)
            if self.early_stopping is not None and self.early_stopping.early_stop:
                break
        return self.history

    # ---- checkpoint: the dictionary of trainer.py:410-417, plus what a resume needs -------------------------------
    def save_checkpoint(self, path: str, epoch: int, rmse: float) -> None:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp"
        torch.save({
            "epoch": epoch,
            "model_state_dict": self.model.state_dict(),
            "optimizer_state_dict": self.optimizer.state_dict(),
            "scheduler_state_dict": self.scheduler.state_dict(),
            "rmse": rmse,
            "history": self.history,
            # additions (ignored by the reference's own loaders, which read model_state_dict only: test/*.py)
            "best_rmse": self.best_rmse, "best_epoch": self.best_epoch,
            "early_stopping": self.early_stopping.state_dict() if self.early_stopping else None,
            "config": asdict(self.config),
        }, tmp)
        os.replace(tmp, path)                        # never leaves a truncated best_model.pth behind

    def load_checkpoint(self, path: str, resume: bool = True) -> dict:
        """``resume=False``: weights only (what the reference's evaluation scripts do).  ``resume=True``: also the
        optimizer moments, scheduler, history, early-stopping counters and the epoch to continue from -- the resume
        path the reference lacks.

        Checkpoints written by the reference load too.  Its Adam was built BEFORE the first forward created
        ``upsample_blocks`` (generator.py:129-130, trainer.py:153-158), so its ``param_groups`` hold only the leading
        parameters of ``model.parameters()`` (the lazy blocks register last): those moments are restored and the
        remaining parameters (trained here when ``optimizer_sees_upsample=True``) start with fresh moments."""
        ck = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(ck["model_state_dict"])       # bumps parameter versions -> packed weights rebuilt
        if resume:
            self._load_optimizer_state(ck["optimizer_state_dict"])
            self.scheduler.load_state_dict(ck["scheduler_state_dict"])
            self.history = ck.get("history", self.history)
            self.start_epoch = int(ck["epoch"]) + 1
            self.best_rmse = float(ck.get("best_rmse", ck.get("rmse", float("inf"))))
            self.best_epoch = int(ck.get("best_epoch", ck["epoch"]))
            if self.early_stopping is not None:
                if ck.get("early_stopping"):
                    self.early_stopping.load_state_dict(ck["early_stopping"])
                elif "rmse" in ck:
                    # reference-written file: it is only ever saved at a new best (trainer.py:386-418), so its rmse IS
                    # the best score so far -- without this the first resumed epoch would overwrite best_model.pth
                    self.early_stopping.best_score = float(ck["rmse"])
                    self.early_stopping.best_epoch = int(ck["epoch"])
        return ck

    def _load_optimizer_state(self, sd: dict) -> None:
        n_saved = sum(len(g["params"]) for g in sd["param_groups"])
        n_have = sum(len(g["params"]) for g in self.optimizer.param_groups)
        if n_saved > n_have:
            raise RuntimeError(
                f"checkpoint optimizer state covers {n_saved} parameters but this trainer optimises {n_have}: it was "
                "written with optimizer_sees_upsample=True; resume it with the same setting")
        if n_saved < n_have:
            if len(sd["param_groups"]) != 1 or len(self.optimizer.param_groups) != 1:
                raise RuntimeError("cannot align a partial optimizer state with several parameter groups")
            sd = {"state": sd["state"], "param_groups": [dict(sd["param_groups"][0])]}
            saved_ids = list(sd["param_groups"][0]["params"])
            fresh = [i for i in range(n_have + len(saved_ids)) if i not in set(saved_ids)][:n_have - n_saved]
            sd["param_groups"][0]["params"] = saved_ids + fresh   # saved ids first: the order matches model.parameters()
        self.optimizer.load_state_dict(sd)


class SyntheticRainBatches:
    """Deterministic stand-in for ``FenheDataset`` + ``DataLoader`` (no dataset ships with the reference): yields
    ``n_batches`` tuples ``(lr [B,T,1,H,W], dem [B,1,Hs,Ws], lu [B,C,Hs,Ws], s_coords [N,2], s_values [B,T,N])`` in
    pinned host memory, the layout of fenhe_dataset.py's samples after collation."""

    def __init__(self, n_batches, B, T, H, W, scale, lu_channels, n_stations=30, seed=0, nan_fraction=0.1):
        g = torch.Generator().manual_seed(seed)
        pin = torch.cuda.is_available()
        self.coords = torch.stack([torch.randint(0, H, (n_stations,), generator=g),
                                   torch.randint(0, W, (n_stations,), generator=g)], 1)
        self.batches = []
        for _ in range(n_batches):
            obs = torch.rand(B, T, n_stations, generator=g) * 20
            obs[torch.rand(B, T, n_stations, generator=g) < nan_fraction] = float("nan")
            items = (torch.rand(B, T, 1, H, W, generator=g) * 5, torch.rand(B, 1, H * scale, W * scale, generator=g),
                     torch.rand(B, lu_channels, H * scale, W * scale, generator=g), self.coords, obs)
            self.batches.append(tuple(t.pin_memory() if pin else t for t in items))

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)
