"""Real multi-GPU check (skips below 2 devices): N ranks over NCCL, one process per GPU, must
reproduce the undivided mesh BIT FOR BIT after a few device-resident cycles -- both through the
C ABI's own transport (ab200_run_cycles_mr: planner, comm stream, grouped ncclSend/ncclRecv and
the dt all-reduce inside libartemis_b200) and through the torch.distributed path.  The loopback
tests in test_gpu_multirank.py cannot see stream-ordering mistakes of the overlapped exchange;
this one can.  The log of each run is kept under gpurun_out/ (copied to profiles/ per round)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("transport", ["native", "torch"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_ranks_on_real_gpus_are_bit_identical_to_the_undivided_mesh(nranks, transport):
    if _ngpus() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    port = 29600 + nranks + (10 if transport == "torch" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "tools", "check_multigpu.py"), "--cycles", "3",
           "--transport", transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"check_multigpu_n{nranks}_{transport}.log"), "w") as fh:
            fh.write(r.stdout[-4000:] + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "BIT-IDENTICAL" in r.stdout


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_ranks_with_sources_and_diffusion_are_bit_identical_to_the_undivided_mesh(nranks):
    """same check with uniform gravity + gas-dust drag + viscosity + conduction configured: every
    stage is split and the diffusion operators read ghost zones the exchange filled."""
    if _ngpus() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + nranks),
           os.path.join(ROOT, "tests", "tools", "check_multigpu.py"), "--cycles", "3",
           "--transport", "native", "--physics"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"check_multigpu_n{nranks}_native_physics.log"), "w") as fh:
            fh.write(r.stdout[-4000:] + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "BIT-IDENTICAL" in r.stdout
