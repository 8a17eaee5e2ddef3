for V in 0 1; do
AQ_DW_CB16=$V timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'depthwise' -c 6 --csv --log-file gpurun_out/r3k_dw_cb16_$V.csv python tools/ncu_decoder.py 64 1 > gpurun_out/r3k_ncu_$V.log 2>&1; echo ncu_decoder rc=$?
done
AQ_DW_CB16=1 timeout 600 python -m pytest tests/test_decoder_gpu.py -m gpu -x -q 2>&1 | tail -2
