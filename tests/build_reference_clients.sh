#!/bin/sh
# Compiles the REFERENCE's own client programs (its unit tests, examples and
# benchmark), unmodified, from where they lie under /root/reference, against THIS
# repository's headers and libraries.  Nothing is copied into the repo; the
# binaries land in tests/_ref_build/ (git-ignored, but they travel to the GPU box).
# This is the drop-in check: code written for afQuantumSim builds and runs on the
# B200 engine.  Usage: tests/build_reference_clients.sh [reference_root]
set -e
REF="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/tests/_ref_build"
INC="-I$ROOT/afquantumsim_b200/host/include -I$ROOT/include"
LIB="-L$ROOT/afquantumsim_b200/lib -lafquantum -laqs_engine -Wl,-rpath,$ROOT/afquantumsim_b200/lib -Wl,-rpath,\$ORIGIN/../../afquantumsim_b200/lib"
mkdir -p "$OUT"
build() { /usr/bin/g++ -std=c++14 -O1 -w $INC "$1" -o "$OUT/$2" $LIB; echo "built $2"; }
build "$REF/test/tests.cpp" ref_tests
for ex in helloworld entanglement superposition fourier_transform grover_search classical_gates \
          classic_2bit_adder draw_circuit quantum_teleportation qft_adder phase_estimation \
          quantum_counting shor_algorithm basis_change; do
    if [ -f "$REF/examples/$ex.cpp" ]; then build "$REF/examples/$ex.cpp" "ex_$ex" || echo "FAILED $ex"; fi
done
build "$REF/benchmark/benchmark.cpp" ref_benchmark || echo "FAILED benchmark"
