mkdir -p gpurun_out
bash scripts/gpu_ci.sh > gpurun_out/ci.out 2>&1; grep -E "exit|passed|failed|FAILED" gpurun_out/ci.out
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_20c.json 2> gpurun_out/bench_20c.err; echo "bench20 rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/bench_20c.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['path_frac'], d['cpu_baseline'], d['e2e']['value'], d['clocks'])"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
