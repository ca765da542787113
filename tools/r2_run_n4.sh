#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus','launches_per_step']}, 'serial', d['serial']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['cpu_baseline'])
t=d['train']; print('train', t['workload'][:40], t['ms_per_step'], t['value'], 'exchange', t['exchange_ms'], t['parallelism'], t['skipped_steps'])
PY
