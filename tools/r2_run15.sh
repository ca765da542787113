#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_audio_gpu.py -x -q -s 2>&1 | tail -40 > gpurun_out/audio_tests.log
cat gpurun_out/audio_tests.log | tail -30
timeout 200 python tools/gl_bench.py 5 --cpu 2>&1 | tail -3 | tee gpurun_out/gl_bench.json
