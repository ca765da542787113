#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_audio_gpu.py -x -q 2>&1 | tail -3
timeout 200 python tools/gl_bench.py 5 2>&1 | tail -2
