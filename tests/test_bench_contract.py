"""bench.py's contract, the parts that run without a GPU: the reference arm (the CPU restatement timed on the host cores)
prints one JSON line with the keys the driver reads; the default arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # as under torchrun: the arm must still use the host's cores
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "deferred+SSAO+SSR Mpixel/s at 4K"
    cb = d["cpu_baseline"]
    # "reference" = the reference's own shader text run by oracle/_ref/libshader_ref.so (when that library is in the tree), "port" = the oracle
    have_text = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "libshader_ref.so"))
    assert cb["kind"] == ("reference" if have_text else "port") and cb["port_value"] > 0
    assert cb["cores"] == len(os.sched_getaffinity(0)) and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
