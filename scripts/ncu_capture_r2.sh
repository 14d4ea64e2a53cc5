#!/bin/bash
# r02: ncu captures of the grouped head kernel (4 batches per launch) + launch list of the bench command.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
python scripts/steady_calls.py voc416_b64 9 4 > gpurun_out/steady_plain.log 2>&1 || { echo "steady_calls failed"; tail -5 gpurun_out/steady_plain.log; exit 1; }
$NCU -k 'regex:head_kernel<\(int\)3' -s 7 -c 1 -o gpurun_out/prof_r02_head_g4 python scripts/steady_calls.py voc416_b64 9 4 > gpurun_out/ncu_r02_head.log 2>&1
$NCU -k 'regex:nms_spec_kernel' -s 7 -c 1 -o gpurun_out/prof_r02_nms_g4 python scripts/steady_calls.py voc416_b64 9 4 > gpurun_out/ncu_r02_nms.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_voc416_b64.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_bench_voc416_b64.csv
