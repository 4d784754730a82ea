#!/bin/bash
# developer probe: one full ncu capture of the persistent blind rotation (3552 gates = two full rounds)
mkdir -p gpurun_out
T=${1:-x}
timeout 300 python tools/time_br.py 3552 > gpurun_out/w12_ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_w12 -s 2 -c 1 -o gpurun_out/prof_w12_$T -f python tools/time_br.py 3552 > gpurun_out/ncu_w12_$T.log 2>&1
echo "rc=$?"; cat gpurun_out/w12_ncu_plain.log
