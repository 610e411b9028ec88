"""aqs_peer_bitswap: the in-place global/local qubit remap over peer memory (tests/peer_cases.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import peer_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_is_an_involution():
    rng = np.random.default_rng(0)
    full = rng.standard_normal(1 << 9).astype(np.float32).view(np.float32)
    once = peer_cases.bitswap_reference(full, 7, [2, 5])
    assert not np.array_equal(once, full)
    assert np.array_equal(peer_cases.bitswap_reference(once, 7, [2, 5]), full)


def test_peer_bitswap_cpu_abi():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "peer_cases.py"), "--abi", "cpu"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "ok peer_cases" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("n_local,local_bits,seed", peer_cases.CASES + [(22, [21, 13, 8], 8), (24, [6], 9)])
def test_peer_bitswap_gpu(n_local, local_bits, seed):
    from afquantumsim_b200 import engine as eng
    eng.ensure_init()
    peer_cases.check_bitswap(eng, n_local, local_bits, seed)
