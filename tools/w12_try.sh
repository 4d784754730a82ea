#!/bin/bash
# developer probe: correctness and device time of the persistent 12-warp blind rotation (variant 70) next to 41 and 60
mkdir -p gpurun_out
{
for v in 70 41 60; do IEACHE_BR_VARIANT=$v timeout 300 python tools/time_br.py 17760; done
IEACHE_BR_VARIANT=70 timeout 300 python tools/time_br.py 1776
IEACHE_BR_VARIANT=70 timeout 300 python tools/time_br.py 3000
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "w12" 2>&1 | tail -5
} > gpurun_out/w12_try.log 2>&1
IEACHE_BR_VARIANT=70 timeout 600 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_w12 -s 1 -c 1 -o gpurun_out/prof_w12_a -f python tools/time_br.py 3552 > gpurun_out/ncu_w12_a.log 2>&1
tail -30 gpurun_out/w12_try.log
