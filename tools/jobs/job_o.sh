for sel in "pointwise:2" "pointwise:4" "pointwise:33" "pointwise:42" "stem:0" "depthwise_tma:1" "depthwise_kernel:1"; do
  k=${sel%%:*}; s=${sel##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/r2b_dec_${k}_${s} python tools/ncu_decoder.py 64 1 > gpurun_out/r2b_ncu_dec_${k}_${s}.log 2>&1; echo "$sel rc=$?"
done
