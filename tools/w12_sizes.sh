#!/bin/bash
# developer probe: W12 (variant 70) device time against launch size, with the latency kernels switched off
mkdir -p gpurun_out
{
for c in 74 148 296 444 592 740 888 1184 1480 1776 2368 3552; do IEACHE_WIDE_MAX=0 IEACHE_CLUSTER_MAX=0 IEACHE_BR_VARIANT=70 timeout 300 python tools/time_br.py $c; done
for c in 592 1184 1776; do IEACHE_WIDE_MAX=0 IEACHE_CLUSTER_MAX=0 IEACHE_BR_VARIANT=41 timeout 300 python tools/time_br.py $c; done
} > gpurun_out/w12_sizes.log 2>&1
cat gpurun_out/w12_sizes.log
