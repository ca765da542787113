#!/bin/bash
# where does the train step's time go: stream / PDL variants + ncu launch list at B4
mkdir -p gpurun_out
export SHAPES=2
echo "== default"; timeout 120 python tools/train_host_time.py 10 2>&1 | tail -2
echo "== no wgrad stream"; VAENAR_NO_WGRAD_STREAM=1 timeout 120 python tools/train_host_time.py 10 2>&1 | tail -2
echo "== no pdl"; VAENAR_NO_PDL=1 timeout 120 python tools/train_host_time.py 10 2>&1 | tail -2
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_b4.csv python tools/ncu_train.py 4 2 > gpurun_out/ncu_train_b4.log 2>&1
echo "ncu rc=$?"
