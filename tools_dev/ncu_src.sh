#!/bin/bash
# Developer tool (GPU box): ncu --set full with source correlation for three launches of one B = 256 forward:
#   the stem (first pw_gemm_tcgen05_v2 launch), the 576-channel depthwise at 15x24, the fused block 2.
# Reports and per-line CSV pages go to gpurun_out/.
set -x
export DBG_N=1
python tools_dev/run_forward.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:pw_gemm_tcgen05_v2 -c 1 -o gpurun_out/ncu_stem -f python tools_dev/run_forward.py > gpurun_out/ncu_stem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_tma_kernel -c 1 -o gpurun_out/ncu_dw576 -f python tools_dev/run_forward.py > gpurun_out/ncu_dw576.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_block_t_kernel -s 1 -c 1 -o gpurun_out/ncu_fbt2 -f python tools_dev/run_forward.py > gpurun_out/ncu_fbt2.log 2>&1
for n in stem dw576 fbt2; do
  ncu -i gpurun_out/ncu_$n.ncu-rep --page raw --csv > gpurun_out/ncu_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$n.ncu-rep --page source --csv > gpurun_out/ncu_${n}_src.csv 2>/dev/null
done
ls -la gpurun_out/ncu_*
