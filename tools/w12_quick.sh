#!/bin/bash
# developer probe: persistent-kernel timing + parity subset + integer-NTT probe
mkdir -p gpurun_out
{
timeout 300 python tools/time_br.py 17760
timeout 300 python tools/time_br.py 35520
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "w12 and (stage or golden or aliasing or exact)" 2>&1 | tail -3
[ -x tools/ntt_probe ] && timeout 120 tools/ntt_probe
} > gpurun_out/w12_quick.log 2>&1
cat gpurun_out/w12_quick.log
