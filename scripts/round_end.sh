#!/bin/bash
# Everything a round leaves behind on one B200: the isolated parity suites, the bench line, the reference arm and the
# ncu launch list of one bench step.  Usage on the GPU box:  bash scripts/round_end.sh   -> files under gpurun_out/
bash scripts/gpu_check.sh > gpurun_out/gpu_check_summary.txt 2>&1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --quick --no-graph --steps 1 --warmup 1 > gpurun_out/b_ncu.log 2>&1
cat gpurun_out/gpu_check_summary.txt
