#!/usr/bin/env python
"""Generate golden vectors from the UNMODIFIED reference (run in the build container).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports ``src.models.convlstm.ConvLSTMCell`` (and ``Generator``) from
``/root/reference`` (override with ``$PLC_REFERENCE``), runs them on CPU in fp32
with fixed seeds, and writes small ``.npz`` fixtures next to this script.  The
reference tree does not travel to the GPU box; these fixtures do.

Fixtures
  cell_*.npz     one ``ConvLSTMCell.forward`` (convlstm.py:16-28) + autograd grads
  rollout_*.npz  the stacked 2-cell T-loop of generator.py:156-171 + autograd grads
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("PLC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
from src.models.convlstm import ConvLSTMCell  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (name, B, Cin, Ch, H, W, k, seed)
CELL_CASES = [
    ("cell_b2_c16_h16_12x15_k3", 2, 16, 16, 12, 15, 3, 101),
    ("cell_b1_c8_h32_9x7_k5", 1, 8, 32, 9, 7, 5, 102),
    ("cell_b2_c16_h32_16x16_k3", 2, 16, 32, 16, 16, 3, 103),
    ("cell_b1_c24_h16_5x6_k3", 1, 24, 16, 5, 6, 3, 104),
    ("cell_b1_c64_h64_16x8_k3", 1, 64, 64, 16, 8, 3, 105),
    ("cell_b3_c1_h16_6x6_k1", 3, 1, 16, 6, 6, 1, 106),
]

# (name, B, T, hidden_dims(2), H, W, seed)  -- wiring of generator.py:57-58
ROLLOUT_CASES = [
    ("rollout_b2_t5_h16_32_8x10", 2, 5, (16, 32), 8, 10, 201),
    ("rollout_b1_t4_h32_16_6x9", 1, 4, (32, 16), 6, 9, 202),
]


def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def gen_cell(name, b, cin, ch, hh, ww, k, seed):
    torch.manual_seed(seed)
    cell = ConvLSTMCell(cin, ch, kernel_size=k)
    x = torch.randn(b, cin, hh, ww, requires_grad=True)
    h = (0.5 * torch.randn(b, ch, hh, ww)).requires_grad_()
    c = torch.randn(b, ch, hh, ww, requires_grad=True)
    h2, c2 = cell(x, h, c)
    gh = torch.randn_like(h2)
    gc = torch.randn_like(c2)
    (h2 * gh).sum().add((c2 * gc).sum()).backward()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        x=np32(x), h=np32(h), c=np32(c),
        weight=np32(cell.conv.weight), bias=np32(cell.conv.bias),
        h_next=np32(h2), c_next=np32(c2), gh=np32(gh), gc=np32(gc),
        dx=np32(x.grad), dh_prev=np32(h.grad), dc_prev=np32(c.grad),
        dW=np32(cell.conv.weight.grad), db=np32(cell.conv.bias.grad),
        k=np.int32(k))


def gen_rollout(name, b, t_steps, hd, hh, ww, seed):
    torch.manual_seed(seed)
    cell1 = ConvLSTMCell(hd[0], hd[0])          # generator.py:57
    cell2 = ConvLSTMCell(hd[0], hd[1])          # generator.py:58
    x_seq = torch.randn(b, t_steps, hd[0], hh, ww, requires_grad=True)
    h1 = torch.zeros(b, hd[0], hh, ww)          # generator.py:156-160
    c1 = torch.zeros_like(h1)
    h2 = torch.zeros(b, hd[1], hh, ww)
    c2 = torch.zeros_like(h2)
    tr = {k_: [] for k_ in ("h1", "c1", "h2", "c2")}
    for t in range(t_steps):                    # generator.py:164-171
        h1, c1 = cell1(x_seq[:, t], h1, c1)
        h2, c2 = cell2(h1, h2, c2)
        for k_, v in (("h1", h1), ("c1", c1), ("h2", h2), ("c2", c2)):
            tr[k_].append(v)
    out = torch.stack(tr["h2"], dim=1)
    d_out = torch.randn_like(out)
    (out * d_out).sum().backward()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        x_seq=np32(x_seq), d_out=np32(d_out),
        w1=np32(cell1.conv.weight), b1=np32(cell1.conv.bias),
        w2=np32(cell2.conv.weight), b2=np32(cell2.conv.bias),
        h1=np32(torch.stack(tr["h1"], 1)), c1=np32(torch.stack(tr["c1"], 1)),
        h2=np32(torch.stack(tr["h2"], 1)), c2=np32(torch.stack(tr["c2"], 1)),
        dx_seq=np32(x_seq.grad),
        dW1=np32(cell1.conv.weight.grad), db1=np32(cell1.conv.bias.grad),
        dW2=np32(cell2.conv.weight.grad), db2=np32(cell2.conv.bias.grad))


def main():
    torch.set_num_threads(1)
    for case in CELL_CASES:
        gen_cell(*case)
    for case in ROLLOUT_CASES:
        gen_rollout(*case)
    print("wrote", len(CELL_CASES) + len(ROLLOUT_CASES), "fixtures to", HERE)


if __name__ == "__main__":
    main()


# ------------------------------------------------------------------------------------------------
# Full Generator + CombinedLoss rollout: state_dict copied AFTER one warm-up forward
# (lazy upsample blocks, generator.py:129-130), non-negative station values (log1p weights).
def gen_generator(name, seed, B, T, H, W, hd, scale, lu_ch, n_st):
    from src.models.generator import Generator
    from src.losses.combined_loss import CombinedLoss
    torch.manual_seed(seed)
    gen = Generator(in_channels=1, dem_channels=1, lu_channels=lu_ch, hidden_dims=list(hd), scale_factor=scale)
    rain = torch.rand(B, T, 1, H, W) * 5.0
    dem = torch.rand(B, 1, H * scale, W * scale)
    lu = torch.rand(B, lu_ch, H * scale, W * scale)
    with torch.no_grad():
        gen(rain, dem, lu)                        # warm-up: creates upsample_blocks
    sd = {k: np32(v) for k, v in gen.state_dict().items()}
    s_coords = torch.stack([torch.randint(0, H, (n_st,)), torch.randint(0, W, (n_st,))], dim=1)
    s_vals = torch.rand(T, n_st) * 20.0
    s_vals[0, 0] = float("nan")                   # missing observation (fenhe_dataset.py keeps NaN for gaps)
    pred = gen(rain, dem, lu)
    loss_mod = CombinedLoss()
    total, parts = loss_mod(pred, rain, s_coords, s_vals, scale_factor=scale)
    total.backward()
    grads = {"grad." + k: np32(p.grad) for k, p in gen.named_parameters() if p.grad is not None}
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        rain=np32(rain), dem=np32(dem), lu=np32(lu), s_coords=s_coords.numpy().astype(np.int64), s_vals=np32(s_vals),
        pred=np32(pred), loss_total=np32(total), loss_point=np32(parts["point"]),
        loss_conserve=np32(parts["conserve"]), loss_smooth=np32(parts["smooth"]),
        loss_temporal=np32(parts["temporal"]), scale=np.int32(scale), hidden_dims=np.array(hd, dtype=np.int32),
        **{"sd." + k: v for k, v in sd.items()}, **grads)


GENERATOR_CASES = [
    ("generator_b2_t3_10x12_h16_32_x2", 301, 2, 3, 10, 12, (16, 32), 2, 5, 7),
    ("generator_b1_t2_8x8_h16_16_x1", 302, 1, 2, 8, 8, (16, 16), 1, 3, 4),
]

if __name__ == "__main__":
    for case in GENERATOR_CASES:
        gen_generator(*case)
    print("wrote", len(GENERATOR_CASES), "generator fixtures")


# ------------------------------------------------------------------------------------------------
# CombinedLoss alone (SURVEY.md section 8f "next-2"): loss terms and d total / d pred from the unmodified reference
# (combined_loss.py:173-191) for every weight strategy, [T,N] and [B,T,N] observations, NaN gaps, stations that
# fall outside the grid, exact ties (sign(0) = 0) and non-default lambdas.
def gen_loss(name, seed, B, T, H, W, scale, n_st, batch_obs, strategy, weighted, lambdas):
    from src.losses.combined_loss import CombinedLoss
    torch.manual_seed(seed)
    pred = (torch.rand(B, T, 1, H * scale, W * scale) * 8.0).requires_grad_(True)
    with torch.no_grad():
        pred[0, 0, 0, 0, :4] = 1.5                           # ties: zero spatial differences
        if T > 1:
            pred[0, 1, 0, 1, :] = pred[0, 0, 0, 1, :]        # ties in time
    lr = torch.rand(B, T, 1, H, W) * 8.0
    coords = torch.stack([torch.randint(0, H, (n_st,)), torch.randint(0, W, (n_st,))], dim=1)
    if n_st > 2:
        coords[1] = torch.tensor([H + 3, 0])                 # outside the grid -> dropped (combined_loss.py:101-107)
        coords[2] = coords[0]                                # two gauges in one pixel
    obs = torch.rand((B, T, n_st) if batch_obs else (T, n_st)) * 60.0
    obs[..., 0, 0] = float("nan")
    if n_st > 3:
        obs[..., -1, 3] = float("nan")
    mod = CombinedLoss(*lambdas, use_weighted_loss=weighted, weight_strategy=strategy)
    total, parts = mod(pred, lr, coords, obs, scale_factor=scale)
    total.backward()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), pred=np32(pred), lr=np32(lr), coords=coords.numpy().astype(np.int64),
        obs=np32(obs), scale=np.int32(scale), lambdas=np.array(lambdas, dtype=np.float32),
        strategy=np.array(strategy), weighted=np.int32(weighted), dpred=np32(pred.grad), total=np32(total),
        point=np32(parts["point"]), conserve=np32(parts["conserve"]), smooth=np32(parts["smooth"]),
        temporal=np32(parts["temporal"]))


LOSS_CASES = [
    ("loss_log_b2_t3_6x7_x2", 401, 2, 3, 6, 7, 2, 6, False, "log", True, (1.0, 1.0, 0.1, 0.05)),
    ("loss_sqrt_b1_t4_5x5_x4_batchobs", 402, 1, 4, 5, 5, 4, 5, True, "sqrt", True, (0.7, 1.3, 0.2, 0.1)),
    ("loss_strat_b3_t2_4x9_x1", 403, 3, 2, 4, 9, 1, 8, True, "stratified", True, (1.0, 1.0, 0.1, 0.05)),
    ("loss_unweighted_b2_t2_8x8_x8", 404, 2, 2, 8, 8, 8, 4, False, "log", False, (2.0, 0.5, 0.3, 0.0)),
]

if __name__ == "__main__":
    for case in LOSS_CASES:
        gen_loss(*case)
    print("wrote", len(LOSS_CASES), "loss fixtures")


# ------------------------------------------------------------------------------------------------
# The reference's optimisation loop (trainer.py:153-158 Adam, :286-315 train_epoch body) on the unmodified Generator
# and CombinedLoss, CPU fp32: per-step loss terms + station RMSE (trainer.py:225-268) and the parameters after N
# steps.  The Trainer class itself cannot be imported here (geopandas / matplotlib), so the loop body is replayed
# statement by statement -- including its quirk: the optimizer is built BEFORE the first forward creates
# upsample_blocks (generator.py:129-130), so those parameters are never updated nor zeroed, yet they are clipped.
def gen_train(name, seed, steps, B, T, H, W, hd, scale, lu_ch, n_st):
    import torch.nn.functional as F
    from src.models.generator import Generator
    from src.losses.combined_loss import CombinedLoss
    torch.manual_seed(seed)
    gen = Generator(in_channels=1, dem_channels=1, lu_channels=lu_ch, hidden_dims=list(hd), scale_factor=scale)
    sd0 = {k: np32(v) for k, v in gen.state_dict().items()}
    opt = torch.optim.Adam(gen.parameters(), lr=5e-4)                         # trainer.py:155-158
    crit = CombinedLoss()
    coords = torch.stack([torch.randint(0, H, (n_st,)), torch.randint(0, W, (n_st,))], dim=1)
    out = {}
    for i in range(steps):
        rain = torch.rand(B, T, 1, H, W) * 5.0
        dem = torch.rand(B, 1, H * scale, W * scale)
        lu = torch.rand(B, lu_ch, H * scale, W * scale)
        obs = torch.rand(B, T, n_st) * 20.0
        obs[0, 0, i % n_st] = float("nan")
        opt.zero_grad()
        fake = gen(rain, dem, lu)
        sf = fake.shape[-2] / rain.shape[-2]
        loss, parts = crit(fake, rain, coords, obs, sf)
        loss.backward()
        if i == 0:
            out.update({"g0." + k: np32(p.grad) for k, p in gen.named_parameters()})      # raw, before the clip
        norm = torch.nn.utils.clip_grad_norm_(gen.parameters(), max_norm=0.5)  # trainer.py:311-314
        out[f"gradnorm{i}"] = np32(norm)
        opt.step()
        if i == 0:
            out.update({"sdA." + k: np32(v) for k, v in gen.state_dict().items()})        # after ONE Adam step
        with torch.no_grad():                                                # trainer.py:225-268
            sc = ((coords.float() + 0.5) * sf - 0.5).long()
            at = fake[:, :, 0][:, :, sc[:, 0], sc[:, 1]]
            m = ~torch.isnan(obs)
            rmse = torch.sqrt(F.mse_loss(at[m], obs[m]))
        out.update({f"rain{i}": np32(rain), f"dem{i}": np32(dem), f"lu{i}": np32(lu), f"obs{i}": np32(obs),
                    f"loss{i}": np.array([float(loss)] + [float(parts[k]) for k in
                                                          ("point", "conserve", "smooth", "temporal")], np.float32),
                    f"rmse{i}": np32(rmse)})
    sd1 = {k: np32(v) for k, v in gen.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), coords=coords.numpy().astype(np.int64),
                        steps=np.int32(steps), scale=np.int32(scale), hidden_dims=np.array(hd, dtype=np.int32),
                        lu_ch=np.int32(lu_ch), **out, **{"sd0." + k: v for k, v in sd0.items()},
                        **{"sd1." + k: v for k, v in sd1.items()})


TRAIN_CASES = [("train_b2_t3_8x10_h16_16_x2_4steps", 501, 4, 2, 3, 8, 10, (16, 16), 2, 3, 6)]

if __name__ == "__main__":
    for case in TRAIN_CASES:
        gen_train(*case)
    print("wrote", len(TRAIN_CASES), "training fixtures")
