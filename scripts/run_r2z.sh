mkdir -p gpurun_out
O=gpurun_out/r2z.out; : > $O
for s in 2 0; do timeout 200 python scripts/tfused_stamps.py $s >> $O 2>&1; done
cat $O
