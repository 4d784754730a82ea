#!/bin/bash
# developer run (GPU box): 4096 request directories through one session call, for several pass sizes
mkdir -p gpurun_out
timeout 1500 python tools/ingest_many.py ${1:-4096} /dev/shm ${2:-64,256,1024,4096} > gpurun_out/ingest_many.log 2>&1; echo "rc=$?"
tail -14 gpurun_out/ingest_many.log
