mkdir -p gpurun_out
O=gpurun_out/r2v.out; : > $O
timeout 300 python -m pytest tests/test_gpu_head.py tests/test_gpu_block.py -q -x -m gpu -k "temporal or clip" >> $O 2>&1
for lib in "" viddet_b200/variants/libviddet_b200_ob1.so; do
 for pf in 0 1; do
  for dbg in 0 1; do
    echo "== lib=$lib prefetch=$pf dbg=$dbg" >> $O
    VD_LIB=$lib VD_TCONV_PREFETCH=$pf VD_TCONV_DBG=$dbg timeout 200 python scripts/tconv_scales.py 2>&1 | grep -v '"dbg": "1".*total' >> $O
  done
 done
done
echo "== L2-resident (8 windows), prefetch 0, mainloop only" >> $O
TC_B=8 VD_TCONV_PREFETCH=0 VD_TCONV_DBG=1 timeout 200 python scripts/tconv_scales.py >> $O 2>&1
cat $O
