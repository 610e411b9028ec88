"""GPU: the same host-layer cases as tests/test_host_cpu.py, in-process, against the
real CUDA engine (no stand-in is ever loaded in this process)."""
import pytest

from tests import host_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", host_cases.ALL, ids=lambda f: f.__name__)
def test_host_layer_on_gpu(case):
    case()
