mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/x_tests.log
timeout 120 python scripts/timeline_probe.py > gpurun_out/x_timeline.txt 2>&1
L2A_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 30 --warmup 3 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
tail -3 gpurun_out/x_tests.log; cut -c1-260 gpurun_out/x_bench.json
