mkdir -p gpurun_out
O=gpurun_out/r2u.out; : > $O
timeout 300 python -m pytest tests/test_gpu_head.py tests/test_gpu_block.py -q -x -m gpu -k "temporal or clip" >> $O 2>&1
echo "== quad off" >> $O
VD_TCONV_QUAD=0 timeout 200 python scripts/tconv_scales.py >> $O 2>&1
echo "== quad on" >> $O
timeout 200 python scripts/tconv_scales.py >> $O 2>&1
echo "== vid t5 bench" >> $O
timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_vid.json 2> gpurun_out/r2u_vid.err
python -c "
import json
for l in open('gpurun_out/r2u_vid.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('vid', round(d['value']), d['ms_per_step'], r['frac'], r['path_frac'], r['kernel_ms'], r['head_kernel_ms'])
" >> $O 2>&1
tail -5 gpurun_out/r2u_vid.err >> $O
cat $O
