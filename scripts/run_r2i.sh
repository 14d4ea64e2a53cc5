mkdir -p gpurun_out
for v in "" s7 s8; do
  if [ -n "$v" ]; then export VD_LIB=$PWD/viddet_b200/variants/libviddet_b200_$v.so; else unset VD_LIB; fi
  for st in 20 2048; do
    python bench.py --steps $st --warmup 5 --no-cpu-baseline > gpurun_out/bench_v${v}_$st.json 2> gpurun_out/bench_v${v}_$st.err; python -c "
import json
d=json.load(open('gpurun_out/bench_v${v}_$st.json')); print('variant [$v] steps $st:', round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['roofline']['path_frac'], d['details']['speculation']['frames_redone_per_step'], d['roofline']['nms_kernel_ms'])"
  done
done
export VD_LIB=$PWD/viddet_b200/variants/libviddet_b200_s7.so
python bench.py --steps 2048 --data video --no-cpu-baseline > gpurun_out/bench_vs7_video.json 2> /dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_vs7_video.json')); print('s7 video:', round(d['value']), d['ms_per_step'], d['details']['speculation'])"
timeout 900 python -m pytest tests/test_gpu_head.py -q -m gpu 2>&1 | tail -2
