mkdir -p gpurun_out
bash scripts/ncu_tfused.sh > gpurun_out/ncu_tfused.out 2>&1; tail -3 gpurun_out/ncu_tfused.out
python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 > gpurun_out/bench_vid416_t5_w64_fused.json 2> gpurun_out/bench_vid416_t5_w64_fused.err; echo "bench vid rc=$?"
VD_TFUSED=0 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vid416_t5_w64_separate.json 2> gpurun_out/bench_vid416_t5_w64_separate.err; echo "bench vid separate rc=$?"
python bench.py --workload vid416_t5_w64 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_vid416_t5_w64_fused_200.json 2> /dev/null; echo "bench vid 200 rc=$?"
timeout 200 python scripts/tfused_stamps.py 2 > gpurun_out/tfused_stamps_s8.txt 2>&1
timeout 200 python scripts/tfused_stamps.py 0 > gpurun_out/tfused_stamps_s32.txt 2>&1
timeout 200 python scripts/tconv_scales.py > gpurun_out/tconv_scales_final.txt 2>&1
python -c "
import json
for f in ('gpurun_out/bench_vid416_t5_w64_fused.json','gpurun_out/bench_vid416_t5_w64_separate.json','gpurun_out/bench_vid416_t5_w64_fused_200.json'):
    d=json.load(open(f)); r=d['roofline']; print(f, d['value'], d['ms_per_step'], r['frac'], r['path_frac'], r['kernel_ms'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
