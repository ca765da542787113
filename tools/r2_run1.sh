#!/bin/bash
# round-2 first GPU check of the fused row kernel: parity test, full GPU suite, bench fused vs per-op chain
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_xblk_gpu.py -x -q -s > gpurun_out/r2_xblk.log 2>&1; echo "xblk rc=$?" 
tail -15 gpurun_out/r2_xblk.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_all.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/r2_gpu_all.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_fused.json 2> gpurun_out/r2_bench_fused.err; echo "bench fused rc=$?"
tail -c 1500 gpurun_out/r2_bench_fused.json
VAENAR_FUSED=0 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_unfused.json 2> gpurun_out/r2_bench_unfused.err; echo "bench unfused rc=$?"
tail -c 600 gpurun_out/r2_bench_unfused.json
