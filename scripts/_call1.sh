mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c1_tests.log
( L2A_TC_FLAGS=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rollout or controller or grbal or tile" 2>&1 | tail -15 ) > gpurun_out/c1_tests_nohint.log
timeout 120 python scripts/mma_rate_probe.py > gpurun_out/c1_mma_rate.txt 2>&1
timeout 120 python scripts/timeline_probe.py > gpurun_out/c1_timeline_hint.txt 2>&1
L2A_TC_FLAGS=1 timeout 120 python scripts/timeline_probe.py > gpurun_out/c1_timeline_nohint.txt 2>&1
L2A_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 30 --warmup 3 > gpurun_out/c1_bench_hint.json 2> gpurun_out/c1_bench_hint.err
L2A_BENCH_SKIP_CPU=1 L2A_TC_FLAGS=1 timeout 200 python bench.py --steps 30 --warmup 3 > gpurun_out/c1_bench_nohint.json 2> gpurun_out/c1_bench_nohint.err
tail -3 gpurun_out/c1_tests.log; cat gpurun_out/c1_bench_hint.json | cut -c1-300
