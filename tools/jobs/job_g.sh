python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_v4.log 2>&1; echo pytest rc=$?; tail -25 gpurun_out/r2_pytest_v4.log | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r2_bench_v3.json 2> gpurun_out/r2_bench_v3.err; echo bench rc=$?; tail -3 gpurun_out/r2_bench_v3.err
python tools/prof_step_ops.py > gpurun_out/r2_step_ops_v3.txt 2>&1; echo prof rc=$?
