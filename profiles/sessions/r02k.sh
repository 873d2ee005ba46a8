#!/bin/bash
# round 2, session k (2 GPUs): the tests that need two devices
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_multi_device.py tests/test_gpu_peer.py -x -q --capture=sys > gpurun_out/r02k_tests.log 2>&1
tail -25 gpurun_out/r02k_tests.log | cut -c1-400
