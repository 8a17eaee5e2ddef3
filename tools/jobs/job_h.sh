python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_v5.log 2>&1; echo pytest rc=$?; tail -25 gpurun_out/r2_pytest_v5.log | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r2_bench_v4.json 2> gpurun_out/r2_bench_v4.err; echo bench rc=$?; tail -3 gpurun_out/r2_bench_v4.err
AQ_WGRAD_BATCH=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r2_bench_v4_wg1.json 2> gpurun_out/r2_bench_v4_wg1.err; echo bench rc=$?
AQ_WGRAD_BATCH=32 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r2_bench_v4_wg32.json 2> gpurun_out/r2_bench_v4_wg32.err; echo bench rc=$?
