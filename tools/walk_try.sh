#!/bin/bash
# developer probe: dependency-driven circuit evaluation (persistent kernel, fused key switch) against level-by-level
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dependency_driven" 2>&1 | tail -5
echo "== level by level"; IEACHE_WALK_MIN=0 timeout 900 python tools/time_circuits.py 5,32,64 5,32,256 4,64,32
echo "== dependency-driven"; timeout 900 python tools/time_circuits.py 5,32,64 5,32,256 4,64,32 5,32,24
} > gpurun_out/walk_try.log 2>&1
cat gpurun_out/walk_try.log
