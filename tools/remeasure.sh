#!/bin/bash
# One GPU call that re-measures the current build: GPU tests, bench line, and (with a second argument) the eager launch
# list + one --set full step.
#   gpurun --timeout 900 -- 'bash tools/remeasure.sh r02c [ncu]'
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.log 2>&1; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-1500
if [ -n "$2" ]; then
NIF_B200_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/step_prof.py 4 > gpurun_out/${TAG}_step_prof.log 2>&1
NIF_B200_GRAPH=0 ncu --set full --clock-control none --import-source on -s 60 -c 40 -f -o gpurun_out/${TAG}_step_full \
    python tools/step_prof.py 4 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_step_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv > gpurun_out/${TAG}_ncu_step.txt 2>&1; cat gpurun_out/${TAG}_ncu_step.txt
python tools/ncu_kernels_json.py gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_ncu_kernels.json $(python -c "import bench; print(bench.BATCH)") > /dev/null 2>&1
fi
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
