#!/bin/bash
# ncu --set full captures of the hot kernels in their steady state (one launch each), run under gpurun AFTER the plain
# programs exited 0.  Reports land in gpurun_out/; summarise here with scripts/ncu_lines.py and copy to profiles/.
# usage: bash scripts/ncu_capture.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
python scripts/steady_calls.py voc416_b64 12 > gpurun_out/steady_plain.log 2>&1 || { echo "steady_calls failed"; exit 1; }
# speculative head kernel = head_kernel<(int)3, ...>; 10th launch = steady state
$NCU -k 'regex:head_kernel<\(int\)3' -s 9 -c 1 -o gpurun_out/prof_${tag}_head python scripts/steady_calls.py voc416_b64 12 > gpurun_out/ncu_${tag}_head.log 2>&1
$NCU -k 'regex:nms_spec_kernel' -s 9 -c 1 -o gpurun_out/prof_${tag}_nms python scripts/steady_calls.py voc416_b64 12 > gpurun_out/ncu_${tag}_nms.log 2>&1
$NCU -k 'regex:head_kernel<\(int\)3' -s 9 -c 1 -o gpurun_out/prof_${tag}_head_coco python scripts/steady_calls.py coco608_b64 12 > gpurun_out/ncu_${tag}_head_coco.log 2>&1
python scripts/steady_block.py > gpurun_out/steady_block_plain.log 2>&1 || { echo "steady_block failed"; exit 1; }
# 2nd forward pass of the block: launches 6..11; the 3x3 expand cells are launches 7, 9, 11 -> skip 7
$NCU -k 'regex:conv_bn_lrelu_kernel' -s 7 -c 1 -o gpurun_out/prof_${tag}_conv3x3 python scripts/steady_block.py > gpurun_out/ncu_${tag}_conv3x3.log 2>&1
$NCU -k 'regex:conv_bn_lrelu_kernel' -s 6 -c 1 -o gpurun_out/prof_${tag}_conv1x1 python scripts/steady_block.py > gpurun_out/ncu_${tag}_conv1x1.log 2>&1
ls -la gpurun_out/prof_${tag}_*.ncu-rep
