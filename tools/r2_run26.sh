#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -x -q > gpurun_out/run26.log 2>&1; tail -60 gpurun_out/run26.log | cut -c1-400
