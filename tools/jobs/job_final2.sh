python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_final.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r2c_pytest_final.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo smoke rc=$?; tail -4 gpurun_out/r2c_smoke.log
python bench.py > gpurun_out/r2c_bench_final.json 2> gpurun_out/r2c_bench_final.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err; echo ref rc=$?
python bench.py --workload decode --images 10000 > gpurun_out/r2c_bench_decode_10k.json 2> gpurun_out/r2c_bench_decode_10k.err; echo decode rc=$?
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c_step_launches.csv python tools/ncu_step.py > gpurun_out/r2c_ncu_step.log 2>&1; echo ncu_step rc=$?
ncu --set full --clock-control none --import-source on -k regex:lora_gemm -c 4 -o gpurun_out/r2c_gemm_final python tools/ncu_gemm_cases.py 1 > gpurun_out/r2c_ncu_gemm_final.log 2>&1; echo ncu_gemm rc=$?
