timeout 900 python -m pytest tests/test_decoder_kernels_gpu.py tests/test_decoder_gpu.py -m gpu -x -q > gpurun_out/r3g_pytest_dec.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r3g_pytest_dec.log | cut -c1-400
for P in 0 128 256; do
AQ_DW_PROMO=$P timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'depthwise|fc_kernel' --csv --log-file gpurun_out/r3g_dw_promo$P.csv python tools/ncu_decoder.py 64 1 > gpurun_out/r3g_ncu_$P.log 2>&1; echo ncu_decoder rc=$?
done
