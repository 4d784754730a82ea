#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/dfma_probe > gpurun_out/dfma_probe.log 2>&1
bash tools/gpu_suite.sh
cat gpurun_out/dfma_probe.log
