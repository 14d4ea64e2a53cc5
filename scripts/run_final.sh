mkdir -p gpurun_out
bash scripts/gpu_ci.sh > gpurun_out/ci.out 2>&1; grep -E "exit|passed|failed|FAILED" gpurun_out/ci.out
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> /dev/null; cut -c1-200 gpurun_out/bench_ref_final.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['path_frac'], d['cpu_baseline']['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']['reasons'])"
bash scripts/ncu_capture_r2.sh > gpurun_out/ncu_final.out 2>&1; tail -3 gpurun_out/ncu_final.out
