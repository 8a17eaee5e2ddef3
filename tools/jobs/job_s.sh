python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_full.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r2b_pytest_full.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo smoke rc=$?; tail -4 gpurun_out/r2b_smoke.log
