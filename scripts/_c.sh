mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/x_tests.log
timeout 300 python scripts/nc_sweep.py 2>/dev/null | head -8 > gpurun_out/x_nc_sweep.jsonl
tail -3 gpurun_out/x_tests.log; cat gpurun_out/x_nc_sweep.jsonl | cut -c1-120
