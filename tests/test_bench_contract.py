"""bench.py contract (CPU part): the reference arm runs without a GPU and prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sequences/s" and d["higher_is_better"] is True
    # default workload = BASELINE configs[2]: the GAN TRAINING step (the metric's first clause), strong scaling
    assert d["metric"] == "gan_train_sequences_per_sec" and d["value"] > 0 and d["scaling"] == "strong"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["global_batch"] == 64 and d["config"]["per_gpu_batch"] == 64
    # both arms must print the SAME config object (the driver compares them): rebuild ours without a GPU
    sys.path.insert(0, ROOT)
    import bench
    ns = type("A", (), {"workload": "train", "no_gan": False, "global_batch": None})()
    assert d["config"] == bench.train_config_dict(ns, 1)


def test_reference_arm_inference_workload_keeps_its_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "infer",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["metric"] == "generator_inference_sequences_per_sec" and d["scaling"] == "weak"
    assert d["config"]["global_batch"] == 32


def test_reference_arm_non_zero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=60, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
