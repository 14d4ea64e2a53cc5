mkdir -p gpurun_out
O=gpurun_out/r2s.out; : > $O
pick() { python -c "
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(sys.argv[2], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],5), 'frac', round(r['frac'],3), 'path', round(r['path_frac'],3), 'kernel_ms', r.get('kernel_ms'), 'head_ms', r.get('head_kernel_ms'), 'redo', d['details']['speculation']['frames_redone_per_step'])
" $1 "$2" >> $O 2>&1; }
echo "== tconv scales" >> $O
timeout 200 python scripts/tconv_scales.py >> $O 2>&1
VD_TCONV_DBG=1 timeout 200 python scripts/tconv_scales.py >> $O 2>&1
VD_TCONV_CTAS=128 timeout 200 python scripts/tconv_scales.py >> $O 2>&1
echo "== vid t5 bench" >> $O
timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_vid_base.json 2> gpurun_out/r2s_vid_base.err; pick gpurun_out/r2s_vid_base.json base
for cfg in "20 128" "16 132" "24 124" "12 136"; do
  set -- $cfg
  VD_HEAD_CTAS=$1 VD_TCONV_CTAS=$2 timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_vid_$1.json 2> gpurun_out/r2s_vid_$1.err; pick gpurun_out/r2s_vid_$1.json "head$1/tconv$2"
done
echo "== voc bench preroll" >> $O
for v in "" "--idle-ms 5" "--preroll-ms 0.3" "--idle-ms 5 --preroll-ms 0.3"; do
  tag=$(echo "$v" | tr -d ' -.')
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $v > gpurun_out/r2s_voc_$tag.json 2> gpurun_out/r2s_voc_$tag.err; pick gpurun_out/r2s_voc_$tag.json "voc [$v]"
done
cat $O
