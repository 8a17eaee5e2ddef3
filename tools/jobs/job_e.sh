python -m pytest tests/test_pretrain_gpu.py tests/test_decoder_gpu.py -m gpu -q > gpurun_out/r2_pytest_v3.log 2>&1; echo pytest rc=$?; tail -60 gpurun_out/r2_pytest_v3.log | cut -c1-300
