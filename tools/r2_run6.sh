#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r2_b_$name.json 2> gpurun_out/r2_b_$name.err; echo "bench $name rc=$?"; }
run single_attn2 VAENAR_SINGLE_CHAIN=1 VAENAR_ATTN=2
run dual_attn2 VAENAR_ATTN=2
run single_attn1 VAENAR_SINGLE_CHAIN=1 VAENAR_ATTN=1
python - <<'PY'
import json
for f in ['single_attn2','dual_attn2','single_attn1']:
    try:
        d=json.loads(open(f'gpurun_out/r2_b_{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, 'ms', round(d['ms_per_step'],4), 'launches', d['launches_per_step'], 'e2e ms', round(d['e2e']['ms_per_step'],4))
PY
