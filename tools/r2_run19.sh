#!/bin/bash
# N = 256 MMAs in the row kernel + two-CTAs-per-SM plain GEMMs: parity first, then A/B timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_run19.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests_run19.log
B="timeout 300 python bench.py --skip-cpu --no-audio --steps 40 --warmup 5 --train-steps 15"
summ() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'inflight ms', round(d['ms_per_step'], 4), 'serial ms', round(d['serial']['ms_per_step'], 4),
      'xrow us', round(d['roofline']['classes']['xblk_row']['ms_per_step'] / 14 * 1e3, 2),
      'train ms', round(d['train']['ms_per_step'], 3) if 'train' in d else None)
PY
}
$B > gpurun_out/ab_default.json 2>gpurun_out/ab_default.err; summ gpurun_out/ab_default.json
VAENAR_XR_N128=1 $B --no-train > gpurun_out/ab_n128.json 2>/dev/null; summ gpurun_out/ab_n128.json
VAENAR_GEMM_OCC2=0 $B > gpurun_out/ab_occ1.json 2>/dev/null; summ gpurun_out/ab_occ1.json
VAENAR_GEMM_OCC2=2 $B > gpurun_out/ab_occ2all.json 2>/dev/null; summ gpurun_out/ab_occ2all.json
$B --no-train --inflight 4 > gpurun_out/ab_inflight4.json 2>/dev/null; summ gpurun_out/ab_inflight4.json
$B --no-train --inflight 2 > gpurun_out/ab_inflight2.json 2>/dev/null; summ gpurun_out/ab_inflight2.json
