#!/bin/bash
# Griffin-Lim: conflict-free shared-memory layout -- parity, timing, ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_audio_gpu.py -x -q 2>&1 | tail -3
timeout 200 python tools/gl_bench.py 5 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k gl_iter_kernel -s 5 -c 1 -o gpurun_out/prof_gl_r2b python tools/gl_bench.py 1 > gpurun_out/ncu_gl.log 2>&1; echo "ncu rc=$?"
