#!/bin/bash
# weight-gradient split-K depth: minimum token blocks per split
mkdir -p gpurun_out
for kb in 1 8 16 24 32; do echo "== min_kb $kb"; VAENAR_WGRAD_MINKB=$kb SHAPES=2 timeout 200 python tools/train_host_time.py 10 2>&1 | tail -2; done
