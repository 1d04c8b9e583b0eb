#!/bin/bash
# usage: scripts/scale_run.sh N   -- multi-rank parity test + headline / cfg5 benches at N GPUs (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -q -k "multi_rank" 2>&1 | tail -3
for c in headline cfg5; do
  NCCL_DEBUG=INFO L2A_BENCH_SKIP_CPU=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N bench.py --gpus $N --steps 20 --warmup 3 --config $c > gpurun_out/r02_bench_${c}_n$N.json 2> gpurun_out/r02_bench_${c}_n$N.err
  grep -c "Init COMPLETE\|comm 0x" gpurun_out/r02_bench_${c}_n$N.err | head -1
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${c}_n$N.json')); print('$c N=$N value %.4g e2e %.4g ms/step %.4f e2e ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_call']))"
done
c=headline
L2A_BENCH_SKIP_CPU=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 296$N bench.py --gpus $N --steps 20 --warmup 3 --config $c --scaling strong > gpurun_out/r02_bench_${c}_strong_n$N.json 2> gpurun_out/r02_bench_${c}_strong_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${c}_strong_n$N.json')); print('$c strong N=$N value %.4g e2e %.4g ms/step %.4f e2e ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_call']))"
