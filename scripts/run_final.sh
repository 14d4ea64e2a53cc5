mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
bash scripts/gpu_ci.sh > gpurun_out/ci.out 2>&1; grep -E "exit|passed|failed|FAILED" gpurun_out/ci.out
VD_LIB=viddet_b200/variants/libviddet_b200_bounds.so timeout 600 python -m pytest tests/test_gpu_head.py tests/test_gpu_block.py -q -m gpu -k "fused_tip or temporal or clip or block or cell" > gpurun_out/bounds_temporal.log 2>&1; tail -2 gpurun_out/bounds_temporal.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> /dev/null; cut -c1-200 gpurun_out/bench_ref_final.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 > gpurun_out/bench_vid416_t5_w64_fused.json 2> gpurun_out/bench_vid416_t5_w64_fused.err; echo "bench vid rc=$?"
python -c "
import json
for f in ('gpurun_out/bench_final.json','gpurun_out/bench_vid416_t5_w64_fused.json'):
    d=json.load(open(f)); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['path_frac'], r.get('traffic'), d['cpu_baseline']['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']['reasons'])"
