#!/bin/bash
# round 2, session j (2 GPUs): device-side reduceDb / re-sort / in-place arenas, full GPU tier
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_reduce_device.py tests/test_gpu_vs_reference.py -x -q > gpurun_out/r02j_reduce.log 2>&1
tail -30 gpurun_out/r02j_reduce.log | cut -c1-300
timeout 1200 python -X faulthandler -m pytest tests -m gpu -x -q --deselect tests/test_gpu_reduce_device.py --deselect tests/test_gpu_vs_reference.py > gpurun_out/r02j_tests.log 2>&1
grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r02j_tests.log | tail -8
grep -n -A12 "Fatal Python" gpurun_out/r02j_tests.log | head -30
