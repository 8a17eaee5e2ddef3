AQ_SILU_NR=1 timeout 900 python -m pytest tests/test_decoder_kernels_gpu.py tests/test_decoder_gpu.py -m gpu -x -q > gpurun_out/r3p_pytest_dec.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r3p_pytest_dec.log | cut -c1-400
for V in 0 1; do
AQ_SILU_NR=$V timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'pointwise' --csv --log-file gpurun_out/r3p_pw_nr$V.csv python tools/ncu_decoder.py 64 1 > gpurun_out/r3p_ncu$V.log 2>&1; echo ncu_decoder rc=$?
done
