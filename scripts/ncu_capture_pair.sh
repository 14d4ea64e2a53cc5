NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
python scripts/steady_block.py > gpurun_out/steady_block_plain.log 2>&1 || exit 1
$NCU -k 'regex:conv_bn_lrelu_pair_kernel' -s 4 -c 1 -o gpurun_out/prof_r1u_conv3x3_pair python scripts/steady_block.py > gpurun_out/ncu_r1u_conv3x3.log 2>&1
python scripts/steady_temporal.py > gpurun_out/steady_temporal_plain.log 2>&1 || exit 1
$NCU -k 'regex:temporal_conv_pair_kernel' -s 3 -c 3 -o gpurun_out/prof_r1u_tconv_pair python scripts/steady_temporal.py > gpurun_out/ncu_r1u_tconv.log 2>&1
ls -la gpurun_out/prof_r1u_*
