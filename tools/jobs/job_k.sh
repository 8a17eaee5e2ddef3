timeout 600 python -m pytest tests/test_lora_gpu.py -m gpu -x -q > gpurun_out/r2b_pytest_lora.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2b_pytest_lora.log | cut -c1-300
timeout 300 python tools/gemm_sweep.py --bn 0 --group 0 > gpurun_out/r2b_sweep_fwd.log 2>&1; echo sweep rc=$?; cat gpurun_out/r2b_sweep_fwd.log
timeout 300 python tools/gemm_sweep.py --bn 0 --group 0 --bwd > gpurun_out/r2b_sweep_bwd.log 2>&1; echo sweepb rc=$?; cat gpurun_out/r2b_sweep_bwd.log
AQ_TRACE_MID=1 AQUALORA_B200_LIB=tools/_trace/libaq_trace.so timeout 300 python tools/gemm_trace.py run "65536,320,320,4096;16384,640,640,1024" > gpurun_out/r2b_trace_mid.log 2>&1; echo trace rc=$?
