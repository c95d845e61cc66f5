#!/bin/bash
# Developer tool (GPU box): the measurement set committed under profiles/ at the end of a round.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --layers 2> gpurun_out/final_layers.txt | tail -1 > gpurun_out/final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final_bench_reference.json
python bench.py --workload decode --steps 20 2>/dev/null | tail -1 > gpurun_out/final_bench_decode.json
python bench.py --workload temporal 2>/dev/null | tail -1 > gpurun_out/final_bench_temporal.json
python bench.py --workload ingest --steps 10 2>/dev/null | tail -1 > gpurun_out/final_bench_ingest.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ls -la gpurun_out/final_*
