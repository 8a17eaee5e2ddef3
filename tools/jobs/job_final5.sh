python -m pytest tests -m gpu -q > gpurun_out/r5_pytest_final.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r5_pytest_final.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5_smoke.log 2>&1; echo smoke rc=$?
python bench.py --workload decode --images 10000 > gpurun_out/r5_bench_decode_10k.json 2> gpurun_out/r5_bench_decode_10k.err; echo decode rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'pointwise|depthwise|stem|se_kernel|fc_kernel|head|expand_dw' --csv --log-file gpurun_out/r5_decoder_launches.csv python tools/ncu_decoder.py 64 1 > gpurun_out/r5_ncu_decoder.log 2>&1; echo ncu_decoder rc=$?
