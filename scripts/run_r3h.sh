mkdir -p gpurun_out
timeout 200 python scripts/tfused_stamps.py 2 > gpurun_out/r3h.out 2>&1
head -22 gpurun_out/r3h.out
