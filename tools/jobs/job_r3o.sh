for U in 1 2 4; do
AQ_DW_ULEN=$U timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'depthwise' --csv --log-file gpurun_out/r3o_dw_u$U.csv python tools/ncu_decoder.py 64 1 > gpurun_out/r3o_ncu$U.log 2>&1; echo ncu_decoder rc=$?
done
