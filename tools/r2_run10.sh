#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_xblk_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/r2_gpu_k.log 2>&1; echo "kernel+model tests rc=$?"; tail -8 gpurun_out/r2_gpu_k.log
timeout 100 python tools/attn_phases.py 2>&1 | tail -16
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --no-train > gpurun_out/r2_bench_a3.json 2> gpurun_out/r2_bench_a3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_a3.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d['roofline']['classes'].items(): print('   ',k, v['launches_per_step'], round(v['ms_per_step'],3), round(v['tflops'],1))
PY
