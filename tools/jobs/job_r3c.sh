timeout 900 python -m pytest tests/test_decoder_kernels_gpu.py tests/test_decoder_gpu.py -m gpu -x -q > gpurun_out/r3c_pytest_dec.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r3c_pytest_dec.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'pointwise|depthwise|stem|se_kernel|fc_kernel|head|expand_dw' --csv --log-file gpurun_out/r3c_decoder_launches.csv python tools/ncu_decoder.py 64 1 > gpurun_out/r3c_ncu_decoder.log 2>&1; echo ncu_decoder rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:depthwise_tma_kernel -s 3 -c 1 -o gpurun_out/r3c_dw3 python tools/ncu_decoder.py 32 1 > gpurun_out/r3c_ncu2.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:depthwise_tma_kernel -s 6 -c 1 -o gpurun_out/r3c_dw5 python tools/ncu_decoder.py 32 1 > gpurun_out/r3c_ncu3.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
