#!/bin/bash
# round-2 late: mel inversion tests + full GPU suite with the reworked SIMT backward kernels + training time
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests_late.log
tail -6 gpurun_out/gpu_tests_late.log
SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
timeout 200 python tools/gl_bench.py 5 2>&1 | tail -1 | tee gpurun_out/gl_bench.json
