#!/bin/bash
# round-2 final evidence: sanitizer on every path, full GPU suite, default bench (+train +mel inversion), c3, reference arm,
# smoke, ncu launch lists of one inference step and one train step
mkdir -p gpurun_out
bash tools/sanitize_all.sh r2 > /dev/null 2>&1; tail -12 gpurun_out/sanitizer_r2.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests_r2.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 > gpurun_out/bench_r2_c3.json 2> gpurun_out/bench_r2_c3.err; echo "bench c3 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; echo "reference arm rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 160 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --skip-cpu --no-train --no-audio --inflight 1 > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_r2.csv python tools/ncu_train.py 32 2 > gpurun_out/ncu_train_r2.log 2>&1; echo "ncu train rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value']/1e6,3), 'ms', round(d['ms_per_step'],4), 'serial', d['serial']['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_ali', d['e2e_with_alignments']['value'])
print('roofline', {k:v for k,v in d['roofline'].items() if k!='classes'})
print('train', {k:v for k,v in d['train'].items() if k not in ('workload','exchange')})
print('mel_inversion', d.get('mel_inversion'))
c=json.loads(open('gpurun_out/bench_r2_c3.json').read().strip().splitlines()[-1])
print('c3', c['ms_per_step'], c['launches_per_step'])
print(open('gpurun_out/bench_r2_reference.json').read()[:400])
PY
