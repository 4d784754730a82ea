#!/bin/bash
# developer probe (GPU box): time every library variant under tools/_variants/ on the same NAND batch, results checked
mkdir -p gpurun_out
cp ie-ache_b200/libieache_b200.so /tmp/lib_shipped.so
{
for f in tools/_variants/lib_*.so; do
  cp $f ie-ache_b200/libieache_b200.so
  echo "== $f"
  timeout 300 python tools/time_br.py ${1:-17760}
  timeout 300 python tools/time_br.py ${1:-17760}
done
cp /tmp/lib_shipped.so ie-ache_b200/libieache_b200.so
} > gpurun_out/w12_ab.log 2>&1
cat gpurun_out/w12_ab.log
