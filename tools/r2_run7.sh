#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_all.log 2>&1; echo "all rc=$?"; tail -15 gpurun_out/r2_gpu_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
r=d.pop('roofline'); t=d.pop('train',None)
print(json.dumps(d,indent=1)[:3500])
print('roofline', {k:v for k,v in r.items() if k!='classes'})
for k,v in r['classes'].items(): print('   ',k, v['launches_per_step'], round(v['ms_per_step'],3), round(v['tflops'],1))
print('train', json.dumps(t,indent=1)[:2500])
PY
