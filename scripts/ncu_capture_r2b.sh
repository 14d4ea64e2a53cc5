#!/bin/bash
# r02: ncu captures of the wide / windowed heads (C = 80 at 608, C = 285 at 416) and the training-side kernels (cfg 5 shapes)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
python scripts/steady_calls.py coco608_b64 6 4 > gpurun_out/steady_plain_coco.log 2>&1 || { echo "steady coco failed"; tail -5 gpurun_out/steady_plain_coco.log; exit 1; }
$NCU -k 'regex:head_kernel<\(int\)3' -s 4 -c 1 -o gpurun_out/prof_r02_head_coco_g4 python scripts/steady_calls.py coco608_b64 6 4 > gpurun_out/ncu_r02_head_coco.log 2>&1
python scripts/steady_calls.py comb416_b64 6 1 > gpurun_out/steady_plain_comb.log 2>&1 || { echo "steady comb failed"; tail -5 gpurun_out/steady_plain_comb.log; exit 1; }
$NCU -k 'regex:head_kernel<\(int\)3' -s 16 -c 4 -o gpurun_out/prof_r02_head_comb python scripts/steady_calls.py comb416_b64 6 1 > gpurun_out/ncu_r02_head_comb.log 2>&1
python scripts/steady_train.py 3 > gpurun_out/steady_plain_train.log 2>&1 || { echo "steady train failed"; tail -5 gpurun_out/steady_plain_train.log; exit 1; }
$NCU -k 'regex:targets_|target_merge|yolo3_loss' -s 10 -c 5 -o gpurun_out/prof_r02_train python scripts/steady_train.py 3 > gpurun_out/ncu_r02_train.log 2>&1
ls -la gpurun_out/prof_r02_*.ncu-rep
