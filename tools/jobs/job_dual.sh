timeout 600 python -m pytest tests/test_lora_gpu.py tests/test_ppft_sd15_gpu.py -m gpu -x -q > gpurun_out/r2d_pytest_lora.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2d_pytest_lora.log | cut -c1-300
SH="65536,1280,320,4096;16384,2560,640,1024;4096,5120,1280,256;4096,1280,1280,256"
for d in 0 1; do
  echo "== AQ_GEMM_DUAL=$d fwd"; AQ_GEMM_DUAL=$d timeout 300 python tools/gemm_sweep.py --shapes "$SH" 2>&1 | tail -4
  echo "== AQ_GEMM_DUAL=$d plain"; AQ_GEMM_DUAL=$d timeout 300 python tools/gemm_sweep.py --plain --shapes "$SH" 2>&1 | tail -4
done
SHB="65536,320,2560,4096;16384,640,5120,1024;4096,1280,10240,256"
for d in 0 1; do echo "== AQ_GEMM_DUAL=$d bwd"; AQ_GEMM_DUAL=$d timeout 300 python tools/gemm_sweep.py --bwd --shapes "$SHB" 2>&1 | tail -3; done
