#!/bin/bash
# ncu evidence for the tensor-core Sobolev step (C4): launch list of eager steps + one --set full capture of its kernels.
#   gpurun --timeout 600 -- 'bash tools/c4_ncu.sh r02v'
TAG=${1:-r02v}
mkdir -p gpurun_out
NIF_B200_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv \
    --log-file gpurun_out/${TAG}_c4_launches.csv python tools/c4_profile.py > gpurun_out/${TAG}_c4_prof.log 2>&1
tail -22 gpurun_out/${TAG}_c4_prof.log
NIF_B200_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'nif_tc_(fwd|bwd_data|bwd_weight|bwd_edge)' -s 24 -c 8 -f \
    -o gpurun_out/${TAG}_c4_full python tools/c4_profile.py > gpurun_out/${TAG}_c4_full.log 2>&1
ncu -i gpurun_out/${TAG}_c4_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_c4_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_c4_raw.csv > gpurun_out/${TAG}_c4_ncu.txt 2>&1; cat gpurun_out/${TAG}_c4_ncu.txt
