"""Drop-in check with the reference's OWN client sources: its unit-test program
(test/tests.cpp, 12 assert-based functions), its examples and its benchmark are
compiled unmodified against this repository's headers and run.

  * CPU (this container, where /root/reference exists): built by
    tests/build_reference_clients.sh and run on the oracle-backed ABI stand-in.
  * GPU box (/root/reference absent): the prebuilt binaries in tests/_ref_build/
    travel with the repo snapshot and run on the real CUDA engine.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "_ref_build")
EXAMPLES = ["helloworld", "entanglement", "superposition", "fourier_transform", "grover_search", "classical_gates",
            "classic_2bit_adder", "draw_circuit", "quantum_teleportation", "qft_adder", "phase_estimation",
            "quantum_counting", "shor_algorithm", "basis_change"]


def run(binary, env_extra=None, timeout=900):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([os.path.join(OUT, binary)], capture_output=True, text=True, timeout=timeout, env=env, cwd=OUT)


@pytest.fixture(scope="module")
def built_on_cpu():
    if not os.path.isdir(REF):
        pytest.skip("reference sources not present on this machine")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    subprocess.check_call([os.path.join(ROOT, "tests", "build_reference_clients.sh"), REF], stdout=subprocess.DEVNULL)
    return {"LD_LIBRARY_PATH": os.path.join(ROOT, "oracle", "_build", "cpu_abi")}


def test_reference_unit_tests_pass_on_cpu_abi(built_on_cpu):
    r = run("ref_tests", built_on_cpu)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "Finished test_special_gates" in r.stdout and "Finished test_qsim_probability" in r.stdout


@pytest.mark.parametrize("name", ["helloworld", "entanglement", "fourier_transform", "grover_search", "draw_circuit",
                                  "quantum_teleportation", "shor_algorithm"])
def test_reference_examples_run_on_cpu_abi(built_on_cpu, name):
    r = run("ex_" + name, built_on_cpu)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    if name == "grover_search":
        assert "Search found <5>" in r.stdout
    if name == "fourier_transform":
        assert "Output measurement after Fourier and Inverse Fourier" in r.stdout


@pytest.mark.gpu
def test_reference_unit_tests_pass_on_gpu():
    if not os.path.exists(os.path.join(OUT, "ref_tests")):
        pytest.skip("tests/_ref_build/ref_tests was not prebuilt")
    r = run("ref_tests")
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "Finished test_special_gates" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", EXAMPLES + ["benchmark"])
def test_reference_examples_run_on_gpu(name):
    binary = "ref_benchmark" if name == "benchmark" else "ex_" + name
    if not os.path.exists(os.path.join(OUT, binary)):
        pytest.skip("not prebuilt")
    r = run(binary)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "sm_100a" in r.stdout or name in ("helloworld", "draw_circuit", "benchmark")
    if name == "grover_search":
        assert "Search found <5>" in r.stdout
