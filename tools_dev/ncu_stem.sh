#!/bin/bash
# Developer tool (GPU box): ncu --set full with source correlation of the stem launch (first pw_gemm_tcgen05_v2 launch) of a B = 256 forward.
set -x
export DBG_N=1
python tools_dev/run_forward.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:pw_gemm_tcgen05_v2 -c 1 -o gpurun_out/ncu_r02_stem -f python tools_dev/run_forward.py > gpurun_out/ncu_r02_stem.log 2>&1
ncu -i gpurun_out/ncu_r02_stem.ncu-rep --page raw --csv > gpurun_out/ncu_r02_stem_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_r02_stem.ncu-rep --page source --csv > gpurun_out/ncu_r02_stem_src.csv 2>/dev/null
