#!/bin/bash
# round 2, session h (1 GPU): full GPU tier with the complete log
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -x -q -v > gpurun_out/r02h_tests.log 2>&1
grep -n "passed\|failed\|error\|Fatal\|PASSED\|FAILED" gpurun_out/r02h_tests.log | tail -15
grep -n -A25 "Fatal Python" gpurun_out/r02h_tests.log | head -60
