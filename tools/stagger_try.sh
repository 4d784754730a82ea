#!/bin/bash
# developer probe: does a per-warp / per-CTA start offset (phase desynchronisation) raise the FP64 pipe's use?
mkdir -p gpurun_out
{
for d in 0 100 250 400 600 900 1500 3000 7000; do IEACHE_W12_STAGGER=$d IEACHE_BR_VARIANT=70 timeout 300 python tools/time_br.py 17760; done
for d in 0 300 700 1200 2000 4000 9000; do IEACHE_BR_STAGGER=$d IEACHE_BR_VARIANT=41 timeout 300 python tools/time_br.py 17760; done
} > gpurun_out/stagger_try.log 2>&1
cat gpurun_out/stagger_try.log
