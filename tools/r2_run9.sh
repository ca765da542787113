#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_edge_cases_gpu.py tests/test_xblk_gpu.py -m gpu -x -q > gpurun_out/r2_gpu_model.log 2>&1; echo "model tests rc=$?"; tail -6 gpurun_out/r2_gpu_model.log
# ncu: full capture of the fused row kernel (decoder variant with alignments = launches 13,14 of a step; prior variant earlier)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xblk_row -s 30 -c 3 -o gpurun_out/prof_xrow_r2 python bench.py --steps 2 --warmup 1 --skip-cpu --no-train --inflight 1 > gpurun_out/ncu_xrow.log 2>&1; echo "ncu xrow rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 110 -c 200 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --skip-cpu --no-train --inflight 1 > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/prof_xrow_r2.ncu-rep
