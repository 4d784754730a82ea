#!/bin/bash
# developer run on an 8-GPU box: a short 8-rank bench (2^17 gates per GPU and step, small expression batches)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --log2-gates 17 --steps 2 --warmup 3 --batch-muladd 64 --batch-mul64 16 --skip-expression --cpu-seconds 2 > gpurun_out/bench_n8_short.json 2> gpurun_out/bench_n8_short.err; echo "bench n8 rc=$?"
tail -c 1200 gpurun_out/bench_n8_short.json; tail -3 gpurun_out/bench_n8_short.err
