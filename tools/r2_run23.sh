#!/bin/bash
# training forward through the fused row kernel (tape outputs) + 128x256 weight-gradient tiles: parity, then A/B timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_run23.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/gpu_tests_run23.log
echo "== default (fused train fwd, wgrad BN 256)"; SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
echo "== wgrad BN 128"; VAENAR_WGRAD_BN=128 SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
echo "== unfused train fwd"; VAENAR_TRAIN_FUSED=0 SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
echo "== unfused train fwd, wgrad BN 128"; VAENAR_TRAIN_FUSED=0 VAENAR_WGRAD_BN=128 SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2
