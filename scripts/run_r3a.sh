mkdir -p gpurun_out
O=gpurun_out/r3a.out; : > $O
timeout 300 python -m pytest tests/test_gpu_head.py tests/test_gpu_block.py -q -x -m gpu -k "fused_tip or temporal or clip or block or cell" >> $O 2>&1
echo "== unfused tip cells (2 staging tiles / 1 staging tile)" >> $O
timeout 200 python scripts/tconv_scales.py >> $O 2>&1
VD_LIB=viddet_b200/variants/libviddet_b200_ob1.so timeout 200 python scripts/tconv_scales.py >> $O 2>&1
echo "== fused per scale" >> $O
timeout 300 python scripts/tfused_scales.py >> $O 2>&1
timeout 200 python scripts/tfused_stamps.py 2 2>&1 | tail -3 >> $O
echo "== vid t5 bench (fused / separate)" >> $O
timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3a_vid.json 2> gpurun_out/r3a_vid.err
VD_TFUSED=0 timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3a_vid_unfused.json 2> gpurun_out/r3a_vid_unfused.err
python -c "
import json
for f in ('gpurun_out/r3a_vid.json','gpurun_out/r3a_vid_unfused.json'):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(f, round(d['value']), d['ms_per_step'], r['frac'], r['path_frac'], r['kernel_ms'], r['head_kernel_ms'], d['details']['speculation']['frames_redone_per_step'])
" >> $O 2>&1
cat $O
