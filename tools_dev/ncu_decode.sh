#!/bin/bash
# Developer tool (GPU box): ncu --set full with source correlation for one launch of the streaming decode kernel.
# $1 = bins per axis, $2 = batch, $3 = tag
mkdir -p gpurun_out
python tools_dev/decode_one.py $1 $2 2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:decode_ori -s 1 -c 1 -o gpurun_out/ncu_$3 -f python tools_dev/decode_one.py $1 $2 2 > gpurun_out/ncu_$3.log 2>&1
ncu -i gpurun_out/ncu_$3.ncu-rep --page raw --csv > gpurun_out/ncu_$3_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$3.ncu-rep --page source --csv > gpurun_out/ncu_$3_src.csv 2>/dev/null
rm -f gpurun_out/ncu_$3.ncu-rep
