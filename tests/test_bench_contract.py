"""bench.py contract pieces that can be checked without a GPU: stdout carries exactly one JSON
line even when a library writes to file descriptor 1 behind Python's back (NCCL prints its
version banner there), the reference arm prints the required keys, rank != 0 of the reference
arm exits quietly, and the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable] + args, capture_output=True, text=True, cwd=ROOT, env=e,
                          timeout=timeout)


def test_stdout_carries_only_the_json_line():
    code = ("import os, sys; sys.path.insert(0, '.'); import bench; bench._claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('stray print'); "
            "bench._emit({'ok': 1})")
    r = _run(["-c", code])
    assert r.returncode == 0
    assert r.stdout == '{"ok": 1}\n'
    assert "NCCL version" in r.stderr and "stray print" in r.stderr


def test_reference_arm_line(oracle_lib):
    r = _run(["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample",
              "32"])
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "zone-cycles/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "zone-cycles/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1"],
             env={"WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["bench.py", "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and r.stdout == ""
    assert "no CPU fallback" in r.stderr
