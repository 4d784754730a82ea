#!/bin/bash
# developer run: the whole GPU test suite, smoke, and a short bench on one B200
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_suite.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/gpu_suite.log 2>&1; echo "smoke rc=$?" >> gpurun_out/gpu_suite.log
timeout 900 python bench.py --log2-gates 17 --steps 2 --warmup 1 --skip-expression --cpu-seconds 2 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench rc=$?" >> gpurun_out/gpu_suite.log
tail -15 gpurun_out/gpu_suite.log; tail -c 1500 gpurun_out/bench_short.json
