bash scripts/gpu_check.sh > gpurun_out/gpu_check_summary.txt 2>&1
python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v7.csv python bench.py --quick --steps 2 --warmup 1 > gpurun_out/b_ncu.log 2>&1
cat gpurun_out/gpu_check_summary.txt
