#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_edge_cases_gpu.py tests/test_xblk_gpu.py -m gpu -x -q > gpurun_out/r2_gpu_model.log 2>&1; echo "model tests rc=$?"; tail -12 gpurun_out/r2_gpu_model.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --no-train > gpurun_out/r2_b_$name.json 2> gpurun_out/r2_b_$name.err; echo "bench $name rc=$?"; tail -2 gpurun_out/r2_b_$name.err; }
run tail X=1
run notail VAENAR_NO_FLOW_TAIL=1
python - <<'PY'
import json
for f in ['tail','notail']:
    try:
        d=json.loads(open(f'gpurun_out/r2_b_{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, 'value', round(d['value']/1e6,3), 'ms', round(d['ms_per_step'],4), 'serial', round(d['serial']['ms_per_step'],4), 'launches', d['launches_per_step'], 'e2e ms', round(d['e2e']['ms_per_step'],4))
PY
