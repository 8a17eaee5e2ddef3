# PDL A/B on the bench probe + tile/group sweep + the new pretrain tests
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v1.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r2_pytest_v1.log
AQ_PDL=1 python tools/gemm_sweep.py --bn 0 --group 0 > gpurun_out/r2_sweep_pdl1.log 2>&1; echo sweep1 rc=$?
AQ_PDL=0 python tools/gemm_sweep.py --bn 0 --group 0 > gpurun_out/r2_sweep_pdl0.log 2>&1; echo sweep0 rc=$?
python tools/gemm_sweep.py --shapes "65536,320,320,4096;16384,640,640,1024;4096,1280,1280,256" --bn 128,160,192 --group 1,2,4 > gpurun_out/r2_sweep_tiles.log 2>&1; echo sweep2 rc=$?
python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_v1_pdl.json 2> gpurun_out/r2_bench_v1.err; echo bench rc=$?
