python tools/sdpa_probe.py > gpurun_out/r2_sdpa_probe.log 2>&1; echo sdpa rc=$?; cat gpurun_out/r2_sdpa_probe.log | cut -c1-400
