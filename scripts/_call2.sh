mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c2_tests.log
timeout 120 python scripts/mma_rate_probe.py > gpurun_out/c2_mma_rate.txt 2>&1
timeout 120 python scripts/timeline_probe.py > gpurun_out/c2_timeline.txt 2>&1
L2A_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 30 --warmup 3 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
L2A_BENCH_SKIP_CPU=1 L2A_NO_GRAPH=1 timeout 200 python bench.py --steps 30 --warmup 3 > gpurun_out/c2_bench_nograph.json 2> gpurun_out/c2_bench_nograph.err
tail -3 gpurun_out/c2_tests.log; cat gpurun_out/c2_bench.json | cut -c1-300
