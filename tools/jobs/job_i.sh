# round 2, session 2: dedicated mid-epilogue warps in lora_gemm (512 threads, setmaxnreg)
timeout 600 python -m pytest tests/test_lora_gpu.py tests/test_deploy_gpu.py -m gpu -x -q > gpurun_out/r2b_pytest_lora.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r2b_pytest_lora.log | cut -c1-300
timeout 300 python tools/gemm_sweep.py --bn 0 --group 0 > gpurun_out/r2b_sweep_fwd.log 2>&1; echo sweep rc=$?; cat gpurun_out/r2b_sweep_fwd.log
timeout 300 python tools/gemm_sweep.py --bn 0 --group 0 --bwd > gpurun_out/r2b_sweep_bwd.log 2>&1; echo sweepb rc=$?; cat gpurun_out/r2b_sweep_bwd.log
timeout 300 python tools/gemm_sweep.py --bn 0 --group 0 --plain > gpurun_out/r2b_sweep_plain.log 2>&1; echo sweepp rc=$?; cat gpurun_out/r2b_sweep_plain.log
AQUALORA_B200_LIB=tools/_trace/libaq_trace.so timeout 300 python tools/gemm_trace.py run "65536,320,320,4096;16384,640,640,1024;65536,1280,320,4096" > gpurun_out/r2b_trace.log 2>&1; echo trace rc=$?
timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-eager-baseline > gpurun_out/r2b_bench_v5.json 2> gpurun_out/r2b_bench_v5.err; echo bench rc=$?; tail -3 gpurun_out/r2b_bench_v5.err
