#!/bin/bash
# Developer tool (GPU box): per-layer table of the first 12 launches for each variant library given on the command line.
mkdir -p gpurun_out/r2
for v in "$@"; do
  SPEF_DEV_LIB=build/var/libspef_$v.so python bench.py --layers --no-cpu-baseline --steps 10 2> gpurun_out/r2/ab_${v}_layers.txt | tail -1 > gpurun_out/r2/ab_${v}_bench.json
  printf "%-10s" $v; head -12 gpurun_out/r2/ab_${v}_layers.txt | awk '{printf "%7.1f", $(NF-5)}'; python -c "import json;d=json.load(open('gpurun_out/r2/ab_${v}_bench.json'));print('  | %.0f img/s' % d['value'])"
done
