#!/bin/bash
# Round measurement on one B200: the bench line, the reference arm, the ncu launch list of the same command, DRAM traffic
# of one blind-rotation launch, and full captures of each shipped kernel (every ncu run only after the same command
# exited 0 without ncu).  Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into profiles/*.
set -x
R=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_${R}_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>/dev/null; echo "ref rc=$?"
BENCH_SHORT="python bench.py --log2-gates 16 --steps 1 --warmup 3 --skip-e2e --skip-expression --cpu-seconds 1 --batch-muladd 0 --batch-mul64 0"
$BENCH_SHORT > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${R}.csv $BENCH_SHORT > gpurun_out/ncu_a.log 2>&1; echo "ncu launches rc=$?"
# DRAM traffic of one 65536-gate blind-rotation launch (the launch size roofline.achieved is computed on)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:blind_rotate_w12 -s 3 -c 1 --csv --log-file gpurun_out/traffic_${R}.csv $BENCH_SHORT > gpurun_out/ncu_t.log 2>&1; echo "ncu traffic rc=$?"
# full captures: persistent blind rotation (3552 gates = 2 rounds), group kernel (592), staged key switch (4096), cluster latency kernel (54), pair (296)
python tools/time_br.py 3552 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_w12 -s 2 -c 1 -o gpurun_out/prof_w12_${R} -f python tools/time_br.py 3552 > gpurun_out/ncu_b.log 2>&1; echo "ncu w12 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:keyswitch_staged -s 2 -c 1 -o gpurun_out/prof_ks_${R} -f python tools/time_br.py 4096 > gpurun_out/ncu_k.log 2>&1; echo "ncu ks rc=$?"
python tools/time_br.py 592 > gpurun_out/plain_g.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 2 -c 1 -o gpurun_out/prof_group_${R} -f python tools/time_br.py 592 > gpurun_out/ncu_g.log 2>&1; echo "ncu group rc=$?"
python tools/time_br.py 54 > gpurun_out/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_cluster -s 2 -c 1 -o gpurun_out/prof_cluster_${R} -f python tools/time_br.py 54 > gpurun_out/ncu_c.log 2>&1; echo "ncu cluster rc=$?"
cat gpurun_out/plain_b.log gpurun_out/plain_g.log gpurun_out/plain_c.log
tail -c 600 gpurun_out/bench_${R}_n1.json
