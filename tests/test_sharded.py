"""Sharded states: world_size-2 (and 4) gloo runs on CPU with the oracle-backed engine
stand-in; the same cases over NCCL on real GPUs when at least two are visible."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(backend, world, port):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dist_cases.py"), "--backend", backend,
                        "--world", str(world), "--port", str(port)], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "ok dist_cases" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_state_gloo_cpu(world):
    run("cpu", world, 29610 + world)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_sharded_state_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    run("cuda", world, 29620 + world)
