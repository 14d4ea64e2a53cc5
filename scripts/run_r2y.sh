mkdir -p gpurun_out
O=gpurun_out/r2y.out; : > $O
timeout 300 python scripts/tfused_scales.py >> $O 2>&1
cat $O
