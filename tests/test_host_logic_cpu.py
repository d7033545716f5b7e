"""Host-side policies that need no GPU: weight-gradient chunking, saved-gates mode parsing, bench batch / traffic
bookkeeping."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_wgrad_chunk_targets_one_million_pixels_per_launch(monkeypatch):
    from plconv import nn as pnn
    monkeypatch.setattr(pnn, "DEFER_WGRAD", "auto")
    # cfg3: 64 sequences of 128x128 per GPU is already 2^20 pixels per step -> the per-step form inside plc_cell_bwd
    assert pnn._wgrad_chunk(10, 64, 128, 128) == 1
    # its 2 / 4 / 8-GPU shards: 2, 4, 8 steps per launch (capped by T)
    assert pnn._wgrad_chunk(10, 32, 128, 128) == 2
    assert pnn._wgrad_chunk(10, 16, 128, 128) == 4
    assert pnn._wgrad_chunk(10, 8, 128, 128) == 8
    assert pnn._wgrad_chunk(3, 8, 128, 128) == 3
    # cfg4: 16 sequences of 256x256 = 2^20 pixels -> per step
    assert pnn._wgrad_chunk(20, 16, 256, 256) == 1
    # the reference's shipped shapes: everything in one launch
    assert pnn._wgrad_chunk(5, 8, 15, 12) == 5
    assert pnn._wgrad_chunk(1, 8, 15, 12) == 1
    monkeypatch.setattr(pnn, "DEFER_WGRAD", "off")
    assert pnn._wgrad_chunk(10, 8, 128, 128) == 1
    monkeypatch.setattr(pnn, "DEFER_WGRAD", "3")
    assert pnn._wgrad_chunk(10, 64, 128, 128) == 3 and pnn._wgrad_chunk(2, 64, 128, 128) == 2


def test_saved_gates_plan_off_needs_no_device(monkeypatch):
    from plconv import nn as pnn
    monkeypatch.setattr(pnn, "SAVE_GATES", "off")
    assert pnn._saved_gates_plan([object(), object()], [None, None], 10, 4, 16, 16, None) == [0, 0]
    assert pnn.LAST_SAVED_GATES_BYTES == 0


def test_layer_streams_disabled_by_default_and_for_one_layer(monkeypatch):
    from plconv import nn as pnn
    assert pnn.LAYER_STREAMS is False or os.environ.get("PLC_LAYER_STREAMS") == "1"
    monkeypatch.setattr(pnn, "LAYER_STREAMS", True)
    assert pnn._layer_streams(None, 1) is None


def test_bench_batch_bookkeeping_and_traffic_scaling():
    import bench
    assert bench.train_batch_sizes(bench.TRAIN, 8) == (64, 8)
    assert bench.train_batch_sizes(bench.RADAR_TRAIN, 8) == (128, 16)
    with pytest.raises(SystemExit):
        bench.train_batch_sizes(dict(bench.TRAIN, global_batch=10), 4)
    full = bench.cfg3_traffic("cell_bwd", True, 64, bench.TRAIN)
    assert full is not None and full > 3e9                                  # 3.57 GB per BPTT call at 64 sequences
    assert bench.cfg3_traffic("cell_bwd", True, 8, bench.TRAIN) == int(full * 8 / 64)
    assert bench.cfg3_traffic("cell_bwd", False, 64, bench.TRAIN) < full    # recompute mode moves fewer bytes
    assert bench.cfg3_traffic("cell_fwd", True, 16, bench.RADAR_TRAIN) is None   # no capture for that shape
