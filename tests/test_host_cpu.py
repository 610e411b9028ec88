"""CPU: the C++ host layer (libafquantum.so) against the reference's known answers
and the oracle, with oracle/_build/cpu_abi/libaqs_engine.so standing in for the CUDA
engine.  Each case runs in a subprocess so the stand-in never shares a process with
any other test (see tests/host_cases.py)."""
import os
import subprocess
import sys

import pytest

from tests import host_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("case", [f.__name__ for f in host_cases.ALL])
def test_host_layer_on_cpu_abi(case):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_cases.py"), "--abi", "cpu", case],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and f"ok {case}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
