"""bench.py contract checks that run without a GPU: the reference arm prints one JSON line with
the agreed keys, and the GPU arm refuses to run (no CPU fallback) when no device is present."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    # --ref-elements is the test hook that shrinks the CPU arm (the driver never passes it: its
    # run covers all 64 Mi elements per step); the printed line must own up to the reduction
    res = run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-elements", str(1 << 20)])
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"] == "fake_quant_fused_fwd_bwd_algorithmic_GBps" and d["value"] > 0
    for key in ("steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["elements"] == 1 << 20 and "REDUCED" in d["cpu_baseline"]["sample"]
    assert d["steps"] == 2 and d["warmup"] == 1 and d["ms_per_step"] > 0


def test_reference_arm_times_the_full_workload_by_default():
    """No extrapolation: ms_per_step is the measured mean of K passes over all 64 Mi elements."""
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "n = args.ref_elements or N_ELEMS" in src and bench.N_ELEMS == 1 << 26
    assert "N_ELEMS / " not in src                                 # the round-1 scale-up of a 1/16 sample is gone


def test_reference_arm_other_ranks_exit_quietly():
    res = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--ref-elements", str(1 << 20)],
              env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    res = run(["--steps", "1", "--warmup", "1", "--train", "none", "--no-cpu"])
    assert res.returncode != 0 and "CUDA" in (res.stderr + res.stdout)


@pytest.mark.gpu
def test_gpu_arm_prints_exactly_one_json_line():
    """stdout of the GPU arm is ONE line (library chatter, e.g. NCCL's version line, goes to stderr) with the contract keys."""
    res = run(["--steps", "5", "--warmup", "3", "--train", "none", "--no-cpu", "--e2e-steps", "2"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = res.stdout.splitlines()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in d, key
    assert d["steps"] == 5 and d["gpu_launches"] == 5 and d["n_gpus"] == 1
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1.2
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * 4 * d["config"]["elements"] == e["d2h_bytes_per_step"]
    assert 0 < e["value"] < e["copy_ceiling"]["value"] * 1.05          # the pipeline cannot beat two plain copies by more than noise
