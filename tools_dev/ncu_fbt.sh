#!/bin/bash
# Developer tool (GPU box): ncu --set full with source correlation for ONE launch of the channel-lane fused kernel of a B = 256
# forward: $1 = how many fused_block_t launches to skip (0 = block 1, 1 = block 2, ...), $2 = tag.
set -x
export DBG_N=1
python tools_dev/run_forward.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:fused_block_t_kernel -s $1 -c 1 -o gpurun_out/ncu_$2 -f python tools_dev/run_forward.py > gpurun_out/ncu_$2.log 2>&1
ncu -i gpurun_out/ncu_$2.ncu-rep --page raw --csv > gpurun_out/ncu_$2_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$2.ncu-rep --page source --csv > gpurun_out/ncu_$2_src.csv 2>/dev/null
ls -la gpurun_out/ncu_$2*
