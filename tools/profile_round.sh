set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01b_n1.json 2> gpurun_out/bench_r01b_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01b_reference.json 2>/dev/null; echo "ref rc=$?"
# launch list (only after the same command exited 0 without ncu)
python bench.py --log2-gates 16 --steps 1 --warmup 3 --skip-e2e --skip-expression --cpu-seconds 1 > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --log2-gates 16 --steps 1 --warmup 3 --skip-e2e --skip-expression --cpu-seconds 1 > gpurun_out/ncu_a.log 2>&1; echo "ncu1 rc=$?"
# full capture of the throughput blind rotation (4096 gates) and of the key switch
python tools/time_br.py 4096 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_kernel -s 2 -c 1 -o gpurun_out/prof_br_r01b -f python tools/time_br.py 4096 > gpurun_out/ncu_b.log 2>&1; echo "ncu2 rc=$?"
# cluster latency kernel, 54 gates
python tools/time_br.py 54 > gpurun_out/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:blind_rotate_cluster -s 2 -c 1 -o gpurun_out/prof_cluster_r01b -f python tools/time_br.py 54 > gpurun_out/ncu_c.log 2>&1; echo "ncu3 rc=$?"
tail -c 600 gpurun_out/bench_r01b_n1.json
