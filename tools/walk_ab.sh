#!/bin/bash
# developer probe (GPU box): circuit batches under every library variant in tools/_variants/
mkdir -p gpurun_out
cp ie-ache_b200/libieache_b200.so /tmp/lib_shipped.so
{
for f in tools/_variants/lib_*.so; do
  cp $f ie-ache_b200/libieache_b200.so
  echo "== $f"
  timeout 600 python tools/time_circuits.py ${@:-5,32,128}
done
cp /tmp/lib_shipped.so ie-ache_b200/libieache_b200.so
} > gpurun_out/walk_ab.log 2>&1
cat gpurun_out/walk_ab.log
