python -m pytest tests -m gpu -x -q -s -k "sd15_ppft_step" > gpurun_out/r2_sd15step.log 2>&1; echo sd15 rc=$? >> gpurun_out/r2_sd15step.log
python bench.py --steps 10 --warmup 3 --shapes-out gpurun_out/r2_shapes_v0.json > gpurun_out/r2_bench_v0.json 2> gpurun_out/r2_bench_v0.err; echo bench rc=$?
ncu --set full --clock-control none --import-source on -k regex:lora_gemm -c 8 -o gpurun_out/r2_gemm_v18 python tools/ncu_gemm_cases.py 1 > gpurun_out/r2_ncu.log 2>&1; echo ncu rc=$?
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v0.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2_pytest_v0.log
