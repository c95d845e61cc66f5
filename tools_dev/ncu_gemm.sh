#!/bin/bash
# Developer tool (GPU box): ncu --set full of the three tail GEMM launches of one B = 256 forward -- L50 (960 -> 320), L51 (320 -> 1280),
# L53 (head, 1280 -> 1736): launches 12, 13, 14 of pw_gemm_tcgen05_v2_kernel (0 = stem, 1..11 = the pointwise layers of blocks 12-17).
set -x
export DBG_N=1
python tools_dev/run_forward.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:pw_gemm_tcgen05_v2 -s 12 -c 3 -o gpurun_out/ncu_r02_gemm -f python tools_dev/run_forward.py > gpurun_out/ncu_r02_gemm.log 2>&1
ncu -i gpurun_out/ncu_r02_gemm.ncu-rep --page raw --csv > gpurun_out/ncu_r02_gemm_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_r02_gemm*
