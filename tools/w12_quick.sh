#!/bin/bash
# developer probe: variant 70 timing + stage parity
mkdir -p gpurun_out
{
IEACHE_BR_VARIANT=${V:-70} timeout 300 python tools/time_br.py 17760
IEACHE_BR_VARIANT=${V:-70} timeout 300 python tools/time_br.py 35520
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "w12 and (stage or golden or aliasing)" 2>&1 | tail -3
} > gpurun_out/w12_quick.log 2>&1
cat gpurun_out/w12_quick.log
