mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
bash scripts/gpu_ci.sh > gpurun_out/ci.out 2>&1; grep -E "exit|passed|failed|FAILED" gpurun_out/ci.out
python __graft_entry__.py --smoke 2>&1 | tail -1
bash scripts/ncu_tfused.sh > gpurun_out/ncu_tfused.out 2>&1; tail -3 gpurun_out/ncu_tfused.out
python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 > gpurun_out/bench_vid416_t5_w64_fused.json 2> gpurun_out/bench_vid416_t5_w64_fused.err; echo "bench vid rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_voc_r3e.json 2> gpurun_out/bench_voc_r3e.err; echo "bench voc rc=$?"
python -c "
import json
for f in ('gpurun_out/bench_vid416_t5_w64_fused.json','gpurun_out/bench_voc_r3e.json'):
    d=json.load(open(f)); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['path_frac'], d['cpu_baseline']['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']['reasons'])"
