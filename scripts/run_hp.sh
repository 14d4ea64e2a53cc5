mkdir -p gpurun_out
O=gpurun_out/hp.out; : > $O
timeout 500 python -m pytest tests/test_gpu_head.py -q -x -m gpu -k "pair_head or fused_tip or temporal or clip" >> $O 2>&1
for v in 1 0; do
VD_HEAD_PAIR=$v timeout 300 python bench.py --workload coco608_b64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/hp_coco_$v.json 2> gpurun_out/hp_coco_$v.err
python -c "
import json
for l in open('gpurun_out/hp_coco_$v.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('coco pair=$v', round(d['value']), d['ms_per_step'], r['frac'], r['path_frac'], r['kernel_ms'], d['details']['speculation']['frames_redone_per_step'])
" >> $O 2>&1
done
timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/hp_vid.json 2> gpurun_out/hp_vid.err
python -c "
import json
for l in open('gpurun_out/hp_vid.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('vid', round(d['value']), d['ms_per_step'], r['frac'], r['path_frac'], r['kernel_ms'])
" >> $O 2>&1
cat $O
