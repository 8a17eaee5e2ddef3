for V in 0 1 2; do
AQ_DW_L2HINT=$V timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'depthwise' --csv --log-file gpurun_out/r3l_dw_hint_$V.csv python tools/ncu_decoder.py 64 1 > gpurun_out/r3l_ncu_$V.log 2>&1; echo ncu_decoder rc=$?
done
