timeout 900 python tools/gemm_sweep.py --bn 0,128,160,192 --group 0,1,2,4,8 > gpurun_out/r2b_sweep_all_fwd.log 2>&1; echo fwd rc=$?
timeout 900 python tools/gemm_sweep.py --plain --bn 0,128,160,192 --group 0,1,2,4,8 > gpurun_out/r2b_sweep_all_plain.log 2>&1; echo plain rc=$?
timeout 900 python tools/gemm_sweep.py --bwd --bn 0,128,160,192 --group 0,1,2,4 > gpurun_out/r2b_sweep_all_bwd.log 2>&1; echo bwd rc=$?
