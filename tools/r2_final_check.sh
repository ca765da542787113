#!/bin/bash
# last sanity of the final tree: full GPU suite
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_r2_final.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests_r2_final.log
