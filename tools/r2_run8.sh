#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_dp_gpu.py -m gpu -q > gpurun_out/r2_gpu_train_dp.log 2>&1; echo "train+dp rc=$?"; tail -8 gpurun_out/r2_gpu_train_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus','launches_per_step']}, d['serial'], d['e2e'], d['cpu_baseline'])
print('train', json.dumps(d['train'],indent=1)[:1800])
PY
