#!/bin/bash
# r02: ncu --set full of the CTA-pair head kernel (COCO-608, 4 batches = 256 frames per launch like the bench), after the plain run exited 0
mkdir -p gpurun_out
python scripts/steady_calls.py coco608_b64 9 4 > gpurun_out/steady_plain_coco.log 2>&1 || { echo "steady_calls failed"; tail -5 gpurun_out/steady_plain_coco.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k 'regex:head_pair_kernel' -s 7 -c 1 -o gpurun_out/prof_r02_hpair_coco python scripts/steady_calls.py coco608_b64 9 4 > gpurun_out/ncu_r02_hpair.log 2>&1
ls -la gpurun_out/prof_r02_hpair_coco.ncu-rep
