#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_session.py tests/test_gpu_e2e_files.py -x -q -m gpu > gpurun_out/gpu_sessions.log 2>&1; echo "rc=$?" >> gpurun_out/gpu_sessions.log
tail -25 gpurun_out/gpu_sessions.log
