bash scripts/gpu_ci.sh
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_20.json 2> gpurun_out/bench_20.err; echo "bench20 rc=$?"; tail -c 1500 gpurun_out/bench_20.err; cut -c1-1500 gpurun_out/bench_20.json
python bench.py --no-cpu-baseline > gpurun_out/bench_2048.json 2> gpurun_out/bench_2048.err; echo "bench2048 rc=$?"; tail -c 600 gpurun_out/bench_2048.err; cut -c1-3000 gpurun_out/bench_2048.json
