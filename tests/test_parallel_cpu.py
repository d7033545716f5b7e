"""World-size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding, bucketed gradient averaging with
.grad views, unused-parameter handling, collective-safe NaN-skip, and the reference-ordered train step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn_name, ret):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import plconv
    from plconv import parallel
    torch.set_num_threads(1)
    r, w, _ = parallel.init_distributed("gloo")
    try:
        ret[rank] = globals()[fn_name](r, w, plconv)
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"worker exited with {p.exitcode}"
    return dict(ret)


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)


def w_grad_average(rank, world, plconv):
    from plconv.parallel import GradReducer, shard_batch
    model = _model()
    x, y = _data()
    sl = shard_batch(8, rank, world)
    red = GradReducer([model[0].parameters(), model[2].parameters()])
    red.zero_grad()
    # per-rank loss = mean over the LOCAL shard; mean over ranks of equal shards == global mean
    ((model(x[sl]) - y[sl]) ** 2).mean().backward()
    red.finish()
    return [p.grad.clone() for p in model.parameters()]


def test_bucketed_grad_average_matches_single_process():
    out = _run("w_grad_average")
    model = _model()
    x, y = _data()
    ((model(x) - y) ** 2).mean().backward()
    for r in (0, 1):
        for g, p in zip(out[r], model.parameters()):
            assert torch.allclose(g, p.grad, atol=1e-6)


def w_unused_and_nan(rank, world, plconv):
    from plconv.parallel import GradReducer, all_ranks_finite
    model = _model()
    extra = torch.nn.Linear(2, 2)                   # never used in forward
    red = GradReducer([model.parameters(), extra.parameters()])
    red.zero_grad()
    x, y = _data()
    model(x).sum().backward()
    red.finish()                                     # must not deadlock on the unused bucket
    ok_all = all_ranks_finite(torch.tensor(1.0))
    bad_one = all_ranks_finite(torch.tensor(float("nan") if rank == 1 else 1.0))
    return ok_all, bad_one, float(extra.weight.grad.abs().sum())


def test_unused_bucket_and_collective_nan_skip():
    out = _run("w_unused_and_nan")
    for r in (0, 1):
        ok_all, bad_one, unused = out[r]
        assert ok_all is True and bad_one is False and unused == 0.0   # both ranks agree to skip


def w_train_step(rank, world, plconv):
    from plconv.parallel import shard_batch
    from plconv.training import TrainStep
    model = _model()
    x, y = _data()
    sl = shard_batch(8, rank, world)
    step = TrainStep(model, [model.parameters()], lr=1e-2, grad_clip_norm=0.5)
    for _ in range(3):
        step(lambda: ((model(x[sl]) - y[sl]) ** 2).mean())
    return [p.detach().clone() for p in model.parameters()]


def test_train_step_matches_single_process_reference_order():
    out = _run("w_train_step")
    model = _model()
    x, y = _data()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for _ in range(3):                                           # trainer.py:290-315 order
        opt.zero_grad()
        ((model(x) - y) ** 2).mean().backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
        opt.step()
    for r in (0, 1):
        for a, b in zip(out[r], model.parameters()):
            assert torch.allclose(a, b, atol=1e-5)


def test_shard_batch():
    from plconv.parallel import shard_batch
    assert [shard_batch(64, r, 8) for r in (0, 7)] == [slice(0, 8), slice(56, 64)]
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)
