for v in g4 g4r88 ""; do
  if [ -n "$v" ]; then export VD_LIB=$PWD/viddet_b200/variants/libviddet_b200_$v.so; else unset VD_LIB; fi
  python bench.py --steps 2048 --no-cpu-baseline > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err; python -c "
import json
d=json.load(open('gpurun_out/bench_var_$v.json')); print('variant [$v]:', d['ms_per_step']*1e3, 'us/step', d['roofline']['kernel_ms_events_around_one_eager_launch'], d['roofline']['nms_kernel_ms'])"
done
