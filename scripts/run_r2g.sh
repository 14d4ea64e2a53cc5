mkdir -p gpurun_out
# own bounds checks compiled in (compute-sanitizer is closed on this pool): the whole GPU suite on the checked library
export VD_LIB=$PWD/viddet_b200/variants/libviddet_b200_bounds.so
( for f in tests/test_gpu_head.py tests/test_gpu_nms.py tests/test_gpu_targets.py tests/test_gpu_edges.py tests/test_gpu_fp32.py tests/test_gpu_train.py; do
    timeout 900 python -m pytest $f -q -m gpu 2>&1 | tail -3
  done ) > gpurun_out/r02_bounds_check.txt 2>&1
grep -c "bounds check failed" gpurun_out/r02_bounds_check.txt; cat gpurun_out/r02_bounds_check.txt | tail -20
unset VD_LIB
python bench.py --workload coco608_b64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_coco_b.json 2> gpurun_out/bench_coco_b.err; python -c "
import json
d=json.load(open('gpurun_out/bench_coco_b.json')); print('coco', d['value'], d['ms_per_step'], d['roofline']['bound'], d['roofline']['frac'], d['roofline']['hbm_frac'], d['roofline']['tensor_frac'], d['roofline']['path_frac'])"
python bench.py --workload comb416_b64 --steps 20 --warmup 5 --group 1 --no-cpu-baseline > gpurun_out/bench_comb.json 2> gpurun_out/bench_comb.err; tail -3 gpurun_out/bench_comb.err; python -c "
import json
d=json.load(open('gpurun_out/bench_comb.json')); print('comb', d['value'], d['ms_per_step'], d['roofline']['bound'], d['roofline']['frac'], d['roofline']['hbm_frac'], d['roofline']['tensor_frac'], d['roofline']['path_frac'])"
python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vidt5_b.json 2> gpurun_out/bench_vidt5_b.err; tail -3 gpurun_out/bench_vidt5_b.err; python -c "
import json
d=json.load(open('gpurun_out/bench_vidt5_b.json')); print('vidt5 clip', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['path_frac'], d['roofline']['kernel_ms'], d['roofline']['head_kernel_ms'], d['roofline']['nms_kernel_ms'], d['e2e'])"
timeout 600 python -m pytest tests/test_gpu_head.py -q -m gpu -k "clip_windows" 2>&1 | tail -3
bash scripts/ncu_capture_r2b.sh
