for v in t125 t135 ""; do
  if [ -n "$v" ]; then export VD_LIB=$PWD/viddet_b200/variants/libviddet_b200_$v.so; else unset VD_LIB; fi
  for data in iid video; do
  python bench.py --steps 2048 --data $data --no-cpu-baseline > gpurun_out/bench_var_${v}_$data.json 2> gpurun_out/bench_var_${v}_$data.err; python -c "
import json
d=json.load(open('gpurun_out/bench_var_${v}_$data.json')); print('variant [$v] $data:', d['ms_per_step']*1e3, 'us/step; redone/step', d['details']['speculation']['frames_redone_per_step'], 'nms', d['roofline']['nms_kernel_ms'])"
  done
done
