#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_xblk_gpu.py -x -q > gpurun_out/r2_xblk.log 2>&1; echo "xblk rc=$?"; tail -3 gpurun_out/r2_xblk.log
timeout 120 python tools/xrow_phases.py > gpurun_out/xrow_phases.log 2>&1; head -100 gpurun_out/xrow_phases.log | grep -vE "hid[0-9] (in|out)"
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r2_b_$name.json 2> gpurun_out/r2_b_$name.err; echo "bench $name rc=$?"; }
run pdl X=1
run single_pdl VAENAR_SINGLE_CHAIN=1
python - <<'PY'
import json
for f in ['pdl','single_pdl']:
    try:
        d=json.loads(open(f'gpurun_out/r2_b_{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, 'ms', round(d['ms_per_step'],4), 'launches', d['launches_per_step'], 'e2e ms', round(d['e2e']['ms_per_step'],4))
PY
