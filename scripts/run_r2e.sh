mkdir -p gpurun_out
bash scripts/gpu_ci.sh > gpurun_out/ci.out 2>&1; grep -E "exit|passed|failed|FAILED" gpurun_out/ci.out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_20b.json 2> gpurun_out/bench_20b.err; echo "bench20 rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/bench_20b.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['path_frac'], d['cpu_baseline'], d['e2e']['value'])"
