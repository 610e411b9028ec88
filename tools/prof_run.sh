set -x
# 1. ncu launch list + full capture of the heaviest specialised pass (default plan: T = 13)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aqs_pass -s 15 -c 15 --csv --log-file gpurun_out/r02_launches.csv python tools/jit_profile.py 30 2 > gpurun_out/r02_prof1.log 2>&1
ncu --set full --clock-control none -k regex:aqs_pass -s 18 -c 1 -o gpurun_out/r02_spec_pass python tools/jit_profile.py 30 2 > gpurun_out/r02_prof2.log 2>&1
ncu --set full --clock-control none -k regex:aqs_pass -s 16 -c 1 -o gpurun_out/r02_spec_pass_light python tools/jit_profile.py 30 2 > gpurun_out/r02_prof3.log 2>&1
# 2. compute-sanitizer on the tile kernels (generic and specialised) and the measurement kernels, small states
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from afquantumsim_b200 import engine as eng, workloads as wl
from tests.test_gpu_engine import random_circuit
from tests.lowering import lower_array
eng.init(0)
import os
for T in (10, 11, 12, 13):
    os.environ["AQS_TILE_BITS"] = str(T)
    for n, seed in ((14, 1), (15, 2), (16, 3)):
        ops = lower_array(random_circuit(n, 150, seed))
        for flags in (eng.PLAN_FUSE, eng.PLAN_FUSE | eng.PLAN_JIT):
            s = eng.State(n); p = eng.Plan(n, ops, flags); s.run(p); s.run(p); s.sync()
            u = np.random.default_rng(0).random(64, dtype=np.float32); s.sample(u); s.close()
print("sanitizer workload done")
PY
compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py > gpurun_out/r02_memcheck.txt 2>&1; tail -5 gpurun_out/r02_memcheck.txt
compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py > gpurun_out/r02_racecheck.txt 2>&1; tail -5 gpurun_out/r02_racecheck.txt
