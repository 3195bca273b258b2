#!/bin/bash
# build_variant.sh NAME "-DFOO=1 ..."  ->  _variants/NAME.so  (load with DODO_LIB=_variants/NAME.so)
# every source is compiled with the extra flags; A/B timing runs compare two builds on the same GPU box
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
mkdir -p $ROOT/_variants/$1
objs=""
for f in common prune pdm hypdist nj align; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr $2 -c $ROOT/dodonaphy_b200/csrc/$f.cu -o $ROOT/_variants/$1/$f.o &
  objs="$objs $ROOT/_variants/$1/$f.o"
done
wait
nvcc -shared -o $ROOT/_variants/$1.so $objs -lcudart
echo $ROOT/_variants/$1.so
