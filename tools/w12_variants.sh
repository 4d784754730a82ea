#!/bin/bash
# developer aid (build container): compile br_w12.cu with each flag combination given on the command line and link one
# library per variant into tools/_variants/ (git-ignored; travels to the GPU box).  tools/w12_ab.sh times them there.
#   usage: tools/w12_variants.sh name1:"-DFLAG=1 ..." name2:"..."
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/ie-ache_b200/csrc
V=$ROOT/tools/_variants
mkdir -p $V
NVCC=/usr/local/cuda/bin/nvcc
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC $flags -c $C/br_w12.cu -o $V/br_w12_$name.o
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $V/lib_$name.so $C/kernels.o $V/br_w12_$name.o $C/engine.o $C/keygen.o $C/circuit.o $C/tfhe_io.o $C/tfhe_compat.o -lcudart_static -lpthread -ldl -lrt
  echo "== $name ($flags)"
  cuobjdump -res-usage $V/br_w12_$name.o | grep -A1 "w12_kernelILi3" | tail -1
  python $ROOT/tools/sass_cost.py $V/br_w12_$name.o "blind_rotate_w12_kernelILi3" auto | grep "dispatch"
done
