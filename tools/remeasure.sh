#!/bin/bash
# First GPU call of the next round: everything that was changed after the last benchmark of round 1 (the 4096-row
# accumulation cap of the tensor-core weight-gradient kernel) re-measured, with the old cap beside it.
#   gpurun --timeout 600 -- 'bash tools/remeasure.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_cap4096.log 2>&1
NIF_B200_TC_WGT_MAX_ROWS=16384 python bench.py > gpurun_out/bench_cap16384.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench_cap4096.log", "gpurun_out/bench_cap16384.log"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["reverse_pass"]["ms"])
PY
NIF_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/launches.csv python tools/step_prof.py 4 > gpurun_out/step_prof.log 2>&1
python tools/big_batch_check.py 22 2>&1 | tail -2
