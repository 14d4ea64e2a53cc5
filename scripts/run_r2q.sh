mkdir -p gpurun_out
for f in tests/test_gpu_head.py tests/test_gpu_fp32.py tests/test_gpu_edges.py tests/test_gpu_ref_exec.py; do timeout 900 python -m pytest $f -q -m gpu 2>&1 | tail -2; done
python scripts/head_only_levels.py 2>&1 | grep level | head -2
for st in 20 2048; do
python bench.py --steps $st --warmup 5 --no-cpu-baseline > gpurun_out/bench_sparse_$st.json 2> gpurun_out/bench_sparse_$st.err; python -c "
import json
d=json.load(open('gpurun_out/bench_sparse_$st.json')); print('steps $st:', round(d['value']), d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['path_frac'], d['details']['speculation']['frames_redone_per_step'])"
done
python bench.py --workload coco608_b64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_coco_sparse.json 2> /dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_coco_sparse.json')); print('coco:', round(d['value']), d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['path_frac'])"
python bench.py --workload vid416_b64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vid_sparse.json 2> /dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_vid_sparse.json')); print('vid:', round(d['value']), d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['path_frac'])"
