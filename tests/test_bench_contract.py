"""bench.py's contract on a machine without a GPU: the reference arm prints one complete JSON line (the CPU reference on its bounded
sample), and the product arm refuses to run instead of falling back to anything on the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"}


def run_bench(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout,
                          env={k: v for k, v in os.environ.items() if k != "AFX_LIB"})


def test_reference_arm_prints_one_complete_line():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built and no /root/reference")
    r = run_bench("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "RANS cell-updates/s" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "synthetic-16M-mixed-omesh"  # the same config as the GPU arm; the sample is named next to it
    assert "synthetic-64k-mixed-omesh" in d["config"]["cpu_arms"] and d["cpu_baseline"]["kind"] == "reference"
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 1e4
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], cwd=ROOT, capture_output=True, text=True,
                       timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_path(afx):
    if afx.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    r = run_bench("--steps", "1", "--warmup", "1", timeout=120)
    assert r.returncode != 0 and "no CUDA device visible" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
