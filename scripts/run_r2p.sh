for pf in 0 1 2 3; do
  VD_HEAD_PREFETCH=$pf python bench.py --steps 2048 --no-cpu-baseline > gpurun_out/bench_pf$pf.json 2> gpurun_out/bench_pf$pf.err; tail -2 gpurun_out/bench_pf$pf.err | cut -c1-200; python -c "
import json
d=json.load(open('gpurun_out/bench_pf$pf.json')); print('prefetch $pf:', d['ms_per_step']*1e3, 'us/step; kernel', d['roofline']['kernel_ms_events_around_one_eager_launch'], 'redone', d['details']['speculation']['frames_redone_per_step'])"
done
VD_HEAD_PREFETCH=2 timeout 600 python -m pytest tests/test_gpu_head.py -q -m gpu -x 2>&1 | tail -2
for pf in 0 2; do
VD_HEAD_PREFETCH=$pf python bench.py --workload coco608_b64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_coco_pf$pf.json 2> /dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_coco_pf$pf.json')); print('coco prefetch $pf:', d['ms_per_step']*1e3, 'us/step', d['roofline']['frac'], d['roofline']['path_frac'])"
done
