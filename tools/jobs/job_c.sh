python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_v2.log 2>&1; echo pytest rc=$?; tail -40 gpurun_out/r2_pytest_v2.log | cut -c1-250
AQ_PDL=0 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_v2_pdl0.json 2> gpurun_out/r2_bench_v2_pdl0.err; echo bench0 rc=$?
AQ_PDL=1 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_v2_pdl1.json 2> gpurun_out/r2_bench_v2_pdl1.err; echo bench1 rc=$?
AQ_PDL=0 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_v2_pdl0b.json 2> gpurun_out/r2_bench_v2_pdl0b.err; echo bench0b rc=$?
