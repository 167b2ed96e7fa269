#!/bin/bash
# hand-field chain kernels: parity tests under a watchdog, then timings
timeout 300 python -m pytest tests/test_gpu_hand16.py -m gpu -q -x -s 2>&1 | grep -vE "Warning|warn" | tail -40 > gpurun_out/hand16_tests.log
echo "exit $?" >> gpurun_out/hand16_tests.log
timeout 200 python tools/prof_hand.py > gpurun_out/prof_hand16.log 2>&1
PROF_HAND_COLOR=0 timeout 200 python tools/prof_hand.py >> gpurun_out/prof_hand16.log 2>&1
cat gpurun_out/hand16_tests.log gpurun_out/prof_hand16.log
