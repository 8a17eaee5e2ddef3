timeout 900 python -m pytest tests/test_decoder_kernels_gpu.py tests/test_decoder_gpu.py -m gpu -x -q > gpurun_out/r4w_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r4w_pytest.log | cut -c1-300
for V in 1 0; do
AQ_PW_WIDE_RES=$V timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'pointwise' -c 16 --csv --log-file gpurun_out/r4w_pw_$V.csv python tools/ncu_decoder.py 64 1 > gpurun_out/r4w_ncu$V.log 2>&1; echo ncu rc=$?
done
