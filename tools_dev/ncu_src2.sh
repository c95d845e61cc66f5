#!/bin/bash
# Developer tool (GPU box): ncu source-level captures of pointwise GEMM launches of one B = 256 forward
# (pw_gemm_tcgen05_v2_kernel: launch 0 = stem, 1 = 96->576 @15x24, 2 = 576->96, 7 = 160->960 @8x12, 8 = 960->160).
export DBG_N=1
python tools_dev/run_forward.py || exit 1
for i in 1 2 7 8; do
  ncu --set full --clock-control none --import-source on -k pw_gemm_tcgen05_v2_kernel -s $i -c 1 -o gpurun_out/ncu_pw$i -f python tools_dev/run_forward.py > gpurun_out/ncu_pw$i.log 2>&1
  ncu -i gpurun_out/ncu_pw$i.ncu-rep --page raw --csv > gpurun_out/ncu_pw${i}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_pw$i.ncu-rep --page source --csv > gpurun_out/ncu_pw${i}_src.csv 2>/dev/null
  rm -f gpurun_out/ncu_pw$i.ncu-rep
done
ls -la gpurun_out/ncu_pw*
