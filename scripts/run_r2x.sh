mkdir -p gpurun_out
O=gpurun_out/r2x.out; : > $O
timeout 300 python -m pytest tests/test_gpu_head.py -q -x -m gpu -k "fused_tip or temporal or clip" >> $O 2>&1
echo "== vid t5 bench (fused)" >> $O
timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x_vid.json 2> gpurun_out/r2x_vid.err
echo "== vid t5 bench (separate kernels)" >> $O
VD_TFUSED=0 timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x_vid_unfused.json 2> gpurun_out/r2x_vid_unfused.err
python -c "
import json
for f in ('gpurun_out/r2x_vid.json','gpurun_out/r2x_vid_unfused.json'):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(f, round(d['value']), d['ms_per_step'], r['frac'], r['path_frac'], r['kernel_ms'], r['head_kernel_ms'], d['details']['speculation'])
" >> $O 2>&1
tail -3 gpurun_out/r2x_vid.err >> $O
cat $O
