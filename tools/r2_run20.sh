#!/bin/bash
# high-priority activation chain in the train step (A/B), batches in flight sweep, ncu of the Griffin-Lim iteration kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_backward_gpu.py tests/test_dp_gpu.py -x -q 2>&1 | tail -3
echo "== prio"; SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
echo "== no prio"; VAENAR_NO_PRIO=1 SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
B="timeout 300 python bench.py --skip-cpu --no-audio --no-train --steps 40 --warmup 5"
for n in 4 5 6 8; do
  $B --inflight $n > gpurun_out/ab_inflight$n.json 2>/dev/null
  python - gpurun_out/ab_inflight$n.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'ms', round(d['ms_per_step'], 4), 'e2e ms', round(d['e2e']['ms_per_step'], 4), 'e2e_ali ms', round(d['e2e_with_alignments']['ms_per_step'], 4))
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k gl_iter_kernel -s 5 -c 1 -o gpurun_out/prof_gl_r2 python tools/gl_bench.py 1 > gpurun_out/ncu_gl.log 2>&1; echo "ncu rc=$?"
