# 2 GPUs: gradient equivalence test, decode sharded over 2 ranks (10k images, all bits checked), PPFT N=2
python -m pytest tests/test_dist_gpu.py tests/test_pretrain_gpu.py -m gpu -q > gpurun_out/r2_pytest_dist_n2.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/r2_pytest_dist_n2.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload decode --images 10000 > gpurun_out/r2_bench_decode_n2.json 2> gpurun_out/r2_bench_decode_n2.err; echo decode2 rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo ppft2 rc=$?
python tools/gemm_sweep.py --shapes "256,320,320,256;256,1280,1280,256" > gpurun_out/r2_sweep_tiny.log 2>&1
python tools/gemm_sweep.py --plain --shapes "256,320,320,256;256,1280,1280,256" >> gpurun_out/r2_sweep_tiny.log 2>&1
