#!/bin/bash
# usage (under gpurun, one GPU): scripts/r02_measure.sh [tests] [bench] [ncu]
# Round-2 evidence run: GPU tests, bench.py for every BASELINE config, the reference arm, ncu launch lists and one full
# capture of the dominant kernel.  Everything lands in gpurun_out/ (copied into profiles/ by hand afterwards).
mkdir -p gpurun_out
WHAT="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/r02_box.txt
if [[ "$WHAT" == *tests* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1
  tail -3 gpurun_out/r02_gpu_tests.log
fi
if [[ "$WHAT" == *bench* ]]; then
  for c in headline cfg1 cfg1p cfg2 cfg3 cfg4 cfg5; do
    timeout 400 python bench.py --steps 20 --warmup 3 --config $c > gpurun_out/r02_bench_${c}_n1.json 2> gpurun_out/r02_bench_${c}_n1.err
    python - <<EOF
import json
try:
    d = json.load(open('gpurun_out/r02_bench_${c}_n1.json'))
    print('$c value %.4g e2e %.4g default %.4g ms/step %.4f kernel_ms %.4f frac %.4f cpu %.4g' % (d['value'], d['e2e']['value'],
          d['e2e']['default_sampler_value'] or 0, d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], (d['cpu_baseline'] or {}).get('value', 0)))
except Exception as ex:
    print('$c FAILED', ex)
EOF
  done
  timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
  cat gpurun_out/r02_bench_reference_arm.json | cut -c1-300
fi
if [[ "$WHAT" == *ncu* ]]; then
  # launch list of the bench command itself (cold, serialised: shares must agree with the bench, not absolutes)
  L2A_BENCH_SKIP_CPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/r02_launches_bench_steps3.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1
  # every host-buffer planning path, one launch per kernel (L2A_NO_GRAPH=1 inside)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
      --log-file gpurun_out/r02_launches_all_paths.csv python scripts/ncu_targets.py all > gpurun_out/r02_ncu_targets.log 2>&1
  # full capture of the dominant kernel (one launch, after the warm-up ones)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_tc2_kernel -s 3 -c 1 \
      -o gpurun_out/r02_rollout_tc2_full -f python scripts/ncu_targets.py headline > gpurun_out/r02_ncu_full.log 2>&1
  ncu -i gpurun_out/r02_rollout_tc2_full.ncu-rep --page raw --csv > gpurun_out/r02_rollout_tc2_full_raw.csv 2>/dev/null
  ls -la gpurun_out | tail -20
fi
