#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_backward_gpu.py tests/test_dp_gpu.py -x -q 2>&1 | tail -6
echo "== lane + attn stream"; SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
echo "== no attn stream"; VAENAR_NO_ATTN_STREAM=1 SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
