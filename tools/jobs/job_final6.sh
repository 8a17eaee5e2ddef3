timeout 400 python bench.py > gpurun_out/r6_bench_final.json 2> gpurun_out/r6_bench_final.err; echo bench rc=$?
