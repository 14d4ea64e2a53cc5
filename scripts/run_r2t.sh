mkdir -p gpurun_out
O=gpurun_out/r2t.out; : > $O
timeout 600 python -m pytest tests/test_gpu_head.py tests/test_gpu_block.py -q -x -m gpu -k "temporal or clip" >> $O 2>&1
echo "== tconv scales (2 staging tiles, 4 stages)" >> $O
timeout 200 python scripts/tconv_scales.py >> $O 2>&1
echo "== tconv scales (1 staging tile, 5 stages)" >> $O
VD_LIB=viddet_b200/variants/libviddet_b200_ob1.so timeout 200 python scripts/tconv_scales.py >> $O 2>&1
echo "== vid t5 bench" >> $O
timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_vid.json 2> gpurun_out/r2t_vid.err
python -c "
import json
for l in open('gpurun_out/r2t_vid.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('vid', round(d['value']), d['ms_per_step'], r['frac'], r['path_frac'], r['kernel_ms'], r['head_kernel_ms'])
" >> $O 2>&1
cat $O
