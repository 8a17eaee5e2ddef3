timeout 900 python -m pytest tests/test_decoder_kernels_gpu.py -m gpu -x -q > gpurun_out/r4t_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r4t_pytest.log | cut -c1-400
