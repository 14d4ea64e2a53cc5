for rep in 1 2; do
for v in prebox ""; do
  if [ -n "$v" ]; then export VD_LIB=$PWD/viddet_b200/variants/libviddet_b200_$v.so; else unset VD_LIB; fi
  python bench.py --steps 2048 --no-cpu-baseline > gpurun_out/bench_ab_$v.json 2> /dev/null; python -c "
import json
d=json.load(open('gpurun_out/bench_ab_$v.json')); print('variant [$v]:', d['ms_per_step']*1e3, 'us/step', d['roofline']['kernel_ms_events_around_one_eager_launch'])"
done
done
