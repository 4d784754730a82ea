#!/bin/bash
# developer run on a 2-GPU box: the two-device test, the bench contract test, and a short 2-rank bench
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_devices or two_contexts or default_kernel_selection" 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_bench_contract.py -x -q -m gpu 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --log2-gates 17 --steps 2 --warmup 3 --batch-muladd 64 --batch-mul64 16 --skip-expression > gpurun_out/bench_n2_short.json 2> gpurun_out/bench_n2_short.err; echo "bench n2 rc=$?"
} > gpurun_out/gpu2_check.log 2>&1
cat gpurun_out/gpu2_check.log; tail -c 2500 gpurun_out/bench_n2_short.json; tail -5 gpurun_out/bench_n2_short.err
