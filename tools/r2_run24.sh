#!/bin/bash
# where the train step's critical path goes: event stamps of the main chain at module boundaries (C3 shape)
mkdir -p gpurun_out
VAENAR_TRAIN_PHASES=1 SHAPES=2 timeout 200 python tools/train_host_time.py 4 2>&1 | grep -E "train phases|B32|B4" | tail -8
echo "== no decoder lane"
VAENAR_NO_DEC_LANE=1 VAENAR_TRAIN_PHASES=1 SHAPES=2 timeout 200 python tools/train_host_time.py 4 2>&1 | grep -E "train phases|B32|B4" | sed -n 5,7p
echo "== no wgrad stream"
VAENAR_NO_WGRAD_STREAM=1 VAENAR_TRAIN_PHASES=1 SHAPES=2 timeout 200 python tools/train_host_time.py 4 2>&1 | grep -E "train phases|B32|B4" | sed -n 5,7p
