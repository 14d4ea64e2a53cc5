#!/bin/bash
# ncu --set full captures of the CTA-pair kernels (after the plain programs exited 0): the 3x3 expand cell of the s16 detection
# block (launch 7 = second cell of the second forward pass) and the three tip-cell launches of one temporal call.
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
python scripts/steady_block.py > gpurun_out/steady_block_plain.log 2>&1 || exit 1
$NCU -k 'regex:conv_bn_lrelu_pair_kernel' -s 7 -c 1 -o gpurun_out/prof_r1u_conv3x3_pair python scripts/steady_block.py > gpurun_out/ncu_r1u_conv3x3.log 2>&1
python scripts/steady_temporal.py > gpurun_out/steady_temporal_plain.log 2>&1 || exit 1
$NCU -k 'regex:temporal_conv_pair_kernel' -s 3 -c 3 -o gpurun_out/prof_r1u_tconv_pair python scripts/steady_temporal.py > gpurun_out/ncu_r1u_tconv.log 2>&1
ls -la gpurun_out/prof_r1u_*
