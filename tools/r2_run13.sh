#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --no-train $EXTRA > gpurun_out/r2_b_$name.json 2> gpurun_out/r2_b_$name.err; echo "bench $name rc=$?"; }
run single X=1
run dual VAENAR_DUAL_CHAIN=1
EXTRA="--inflight 3" run single3 X=1
EXTRA="--inflight 1" run dual1 VAENAR_DUAL_CHAIN=1
python - <<'PY'
import json
for f in ['single','dual','single3','dual1']:
    try:
        d=json.loads(open(f'gpurun_out/r2_b_{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, 'value', round(d['value']/1e6,3), 'ms', round(d['ms_per_step'],4), 'serial', round(d['serial']['ms_per_step'],4), 'launches', d['launches_per_step'], 'e2e ms', round(d['e2e']['ms_per_step'],4), 'e2e_ali', round(d['e2e_with_alignments']['ms_per_step'],4))
PY
