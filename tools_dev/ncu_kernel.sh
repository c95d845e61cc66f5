#!/bin/bash
# Developer tool (GPU box): ncu --set full with source correlation for ONE launch of a kernel of a B = 256 forward:
# $1 = kernel-name regex, $2 = how many matching launches to skip, $3 = tag.
set -x
export DBG_N=1
python tools_dev/run_forward.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/ncu_$3 -f python tools_dev/run_forward.py > gpurun_out/ncu_$3.log 2>&1
ncu -i gpurun_out/ncu_$3.ncu-rep --page raw --csv > gpurun_out/ncu_$3_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$3.ncu-rep --page source --csv > gpurun_out/ncu_$3_src.csv 2>/dev/null
rm -f gpurun_out/ncu_$3.ncu-rep
ls -la gpurun_out/ncu_$3*
