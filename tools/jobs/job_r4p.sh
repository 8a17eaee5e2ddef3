timeout 600 python tools/prof_adds.py > gpurun_out/r4_prof_adds.log 2>&1; echo rc=$?
