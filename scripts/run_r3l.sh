mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --kernel-name-base demangled -k 'regex:temporal_head_fused' -s 4 -c 3 --csv --log-file gpurun_out/r02_ncu_tfused_clip_w64_traffic.csv python scripts/tfused_scales.py > gpurun_out/ncu_tfused_clip.log 2>&1
cat gpurun_out/r02_ncu_tfused_clip_w64_traffic.csv | tail -20
