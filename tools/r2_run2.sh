#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_xblk_gpu.py -x -q -s > gpurun_out/r2_xblk.log 2>&1; echo "xblk rc=$?"
tail -9 gpurun_out/r2_xblk.log
timeout 120 python tools/xrow_phases.py > gpurun_out/xrow_phases.log 2>&1; head -90 gpurun_out/xrow_phases.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r2_bench_fused.json 2> gpurun_out/r2_bench_fused.err; echo "bench fused rc=$?"
VAENAR_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r2_bench_fused_nopdl.json 2> gpurun_out/r2_bench_fused_nopdl.err; echo "bench fused nopdl rc=$?"
python - <<'PY'
import json
for f in ['r2_bench_fused','r2_bench_fused_nopdl']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    print(f, round(d['ms_per_step'],4), d['launches_per_step'], round(d['e2e']['ms_per_step'],4))
    for k,v in d['roofline']['classes'].items():
        print('   ',k, v['launches_per_step'], round(v['ms_per_step'],3), round(v['ms_per_step']/v['launches_per_step']*1e3,1),'us', round(v['tflops'],1))
PY
