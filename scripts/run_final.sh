mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
python -m pytest tests/ -x -q -m gpu > gpurun_out/gpu_tests_one_process.log 2>&1; tail -2 gpurun_out/gpu_tests_one_process.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> /dev/null; cut -c1-160 gpurun_out/bench_ref_final.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_final.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], r['frac'], r['path_frac'], d['cpu_baseline']['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']['reasons'])"
bash scripts/ncu_hpair.sh > gpurun_out/ncu_hpair.out 2>&1; tail -1 gpurun_out/ncu_hpair.out
