#!/bin/bash
# weight-gradient epilogue: 8 warps + 16-byte vector reductions
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py tests/test_backward_gpu.py -x -q > gpurun_out/run28.log 2>&1; tail -30 gpurun_out/run28.log | cut -c1-400
echo "== default"; SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
