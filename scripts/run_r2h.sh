mkdir -p gpurun_out
for f in tests/test_gpu_targets.py tests/test_gpu_train.py tests/test_gpu_ref_exec.py; do timeout 600 python -m pytest $f -q -m gpu 2>&1 | tail -2; done
python bench.py --workload targets_c285_b128 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_targets_b.json 2> gpurun_out/bench_targets_b.err; tail -2 gpurun_out/bench_targets_b.err; python -c "
import json
d=json.load(open('gpurun_out/bench_targets_b.json')); print('targets bulk', d['value'], d['ms_per_step'], d['roofline']['frac'])"
VD_TARGETS_FILL=0 python bench.py --workload targets_c285_b128 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_targets_c.json 2> gpurun_out/bench_targets_c.err; python -c "
import json
d=json.load(open('gpurun_out/bench_targets_c.json')); print('targets stcs', d['value'], d['ms_per_step'], d['roofline']['frac'])"
python scripts/bench_configs.py train 2>/dev/null | tail -3
