#!/bin/bash
# ncu --set full of the three tip-cell launches (cfg 4 shapes, one per scale), after the plain run has exited 0
mkdir -p gpurun_out
timeout 200 python scripts/tconv_scales.py > gpurun_out/tconv_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/tconv_plain.log; exit 1; }
cat gpurun_out/tconv_plain.log
# launches per scale: 3 warm + 20 timed = 23; capture the 10th launch of every scale (skip 9, then every 23rd)
for i in 0 1 2; do
  skip=$((9 + 23 * i))
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k 'regex:temporal_conv' -s $skip -c 1 -o gpurun_out/prof_r02_tconv_s$i python scripts/tconv_scales.py > gpurun_out/ncu_tconv_$i.log 2>&1
done
ls -la gpurun_out/prof_r02_tconv_s*.ncu-rep
