#!/bin/bash
O=gpurun_out; mkdir -p $O
touch tests/__init__.py
CUDA_LAUNCH_BLOCKING=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python profiles/experiments/brick_debug.py 32 100 > $O/r2_v8_sanitizer.log 2>&1
grep -v "^=========     at\|^=========         \[" $O/r2_v8_sanitizer.log | head -60
