#!/bin/bash
# round-2 evidence run: full GPU test suite, default bench, ncu launch list + full captures of the two hot kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gpu_tests_r2.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; echo "reference arm rc=$?"
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 > gpurun_out/bench_r2_c3.json 2> gpurun_out/bench_r2_c3.err; echo "bench c3 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 160 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --skip-cpu --no-train --inflight 1 > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xblk_row -s 24 -c 4 -o gpurun_out/prof_xrow_r2 -f python bench.py --steps 2 --warmup 1 --skip-cpu --no-train --inflight 1 > gpurun_out/ncu_xrow.log 2>&1; echo "ncu xrow rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention3 -s 24 -c 2 -o gpurun_out/prof_attn3_r2 -f python bench.py --steps 2 --warmup 1 --skip-cpu --no-train --inflight 1 > gpurun_out/ncu_attn3.log 2>&1; echo "ncu attn3 rc=$?"
timeout 120 python tools/xrow_phases.py > gpurun_out/xrow_phases.log 2>&1; grep -E "kernel span" gpurun_out/xrow_phases.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1.json').read().strip().splitlines()[-1])
print('value', round(d['value']/1e6,3), 'ms', round(d['ms_per_step'],4), 'serial', d['serial'], 'e2e', d['e2e'], 'e2e_ali', d['e2e_with_alignments'])
print('roofline', {k:v for k,v in d['roofline'].items() if k!='classes'})
for k,v in d['roofline']['classes'].items(): print('   ',k, v['launches_per_step'], round(v['ms_per_step'],3), round(v['tflops'],1))
print('standin', d['gpu_eager_standin']); print('cpu', d['cpu_baseline']); print('train', {k:v for k,v in d['train'].items() if k not in ('workload','exchange')})
print(open('gpurun_out/bench_r2_reference.json').read()[:600])
PY
