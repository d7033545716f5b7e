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


class _FakeCell(torch.nn.Module):
    """Stands in for plconv.nn.ConvLSTMCell on CPU: a `conv` parameter holder plus the `_packed` marker attribute; the
    fused rollout's backward hands (dW, db) to `_grad_sink` when the layer's t = 0 step has been queued."""

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 8, 3, padding=1)

    def _packed(self, need_dgrad):
        raise AssertionError("never called on CPU")


def w_grad_sink(rank, world, plconv):
    from plconv.parallel import GradReducer, nonfinite_flag
    torch.manual_seed(0)
    top, bottom, other = _FakeCell(), _FakeCell(), torch.nn.Linear(2, 2)
    red = GradReducer([top.parameters(), list(bottom.parameters()) + list(other.parameters())])
    assert red.attach_cell_sinks(torch.nn.ModuleList([top, bottom, other])) == 2
    red.zero_grad()
    g = torch.Generator().manual_seed(10 + rank)
    gw = [torch.randn(8, 3, 3, 3, generator=g) for _ in range(2)]
    gb = [torch.randn(8, generator=g) for _ in range(2)]
    # what _StackRolloutFn.backward does at t = 0, top layer first
    assert top._grad_sink(top, gw[0], gb[0]) is True
    launched_top = red.buckets[0]["pending"] == 0 and (world == 1 or red.buckets[0]["handle"] is not None)
    assert bottom._grad_sink(bottom, gw[1], gb[1]) is True
    launched_bottom_early = red.buckets[1]["pending"] == 0       # `other` has not reported yet: must still be pending
    other(torch.ones(1, 2)).sum().backward()                      # ordinary autograd hook completes the mixed bucket
    red.finish()
    flag = nonfinite_flag(torch.tensor(float("inf") if rank == 1 else 0.0))
    red.remove()
    return (top.conv.weight.grad.clone(), bottom.conv.bias.grad.clone(), launched_top, launched_bottom_early,
            float(flag), top._grad_sink is None)


def test_per_layer_gradient_sink_and_device_flag():
    out = _run("w_grad_sink")
    gens = [torch.Generator().manual_seed(10 + r) for r in (0, 1)]
    gw = [[torch.randn(8, 3, 3, 3, generator=g) for _ in range(2)] for g in gens]
    gb = [[torch.randn(8, generator=g) for _ in range(2)] for g in gens]
    for r in (0, 1):
        w_top, b_bottom, launched_top, early, flag, removed = out[r]
        assert torch.allclose(w_top, (gw[0][0] + gw[1][0]) / 2, atol=1e-6)
        assert torch.allclose(b_bottom, (gb[0][1] + gb[1][1]) / 2, atol=1e-6)
        assert launched_top and not early and flag == 1.0 and removed
