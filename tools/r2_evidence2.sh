#!/bin/bash
# round-2 late evidence: sanitizer on every path (incl. mel inversion), full GPU suite, default bench (+train +mel inversion), c3
mkdir -p gpurun_out
bash tools/sanitize_all.sh r2 > /dev/null 2>&1; tail -12 gpurun_out/sanitizer_r2.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests_r2.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 > gpurun_out/bench_r2_c3.json 2> gpurun_out/bench_r2_c3.err; echo "bench c3 rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value']/1e6,3), 'ms', round(d['ms_per_step'],4), 'serial', d['serial']['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_ali', d['e2e_with_alignments']['value'])
print('train', {k:v for k,v in d['train'].items() if k not in ('workload','exchange')})
print('mel_inversion', d.get('mel_inversion'))
c=json.loads(open('gpurun_out/bench_r2_c3.json').read().strip().splitlines()[-1])
print('c3', c['ms_per_step'], c['launches_per_step'])
PY
