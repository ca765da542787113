#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -x -q > gpurun_out/dp_test_r2.log 2>&1; echo "dp rc=$?"; tail -4 gpurun_out/dp_test_r2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus','launches_per_step']}, d['serial']['ms_per_step'], d['e2e']['ms_per_step'], d['cpu_baseline'])
t=d['train']; print('train', t['ms_per_step'], t['value'], t['exchange_ms'], t['parallelism'], t['skipped_steps'])
PY
