#!/bin/bash
# r02: ncu --set full of the fused temporal head kernel (cfg 4 shapes: 32 windows of T=5, C=30) + launch list of the cfg 4 bench command.
mkdir -p gpurun_out
python scripts/steady_temporal.py > gpurun_out/steady_temporal_plain.log 2>&1 || { echo "steady_temporal failed"; tail -5 gpurun_out/steady_temporal_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k 'regex:temporal_head_fused' -s 2 -c 1 -o gpurun_out/prof_r02_tfused python scripts/steady_temporal.py > gpurun_out/ncu_r02_tfused.log 2>&1
python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vidt5_plain.json 2> gpurun_out/bench_vidt5_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_vid416_t5_w64.csv python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench_vidt5.log 2>&1
ls -la gpurun_out/prof_r02_tfused.ncu-rep gpurun_out/r02_launches_bench_vid416_t5_w64.csv
