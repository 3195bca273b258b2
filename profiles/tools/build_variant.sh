#!/bin/bash
# build_variant.sh NAME "-DFOO=1 ..."  ->  _variants/NAME.so  (load with DODO_LIB=_variants/NAME.so)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
mkdir -p $ROOT/_variants/$1
for f in common prune pdm nj; do
  if [ "$f" = "prune" ] || [ ! -f $ROOT/_variants/base_$f.o ]; then
    out=$ROOT/_variants/$1/$f.o
    [ "$f" != "prune" ] && out=$ROOT/_variants/base_$f.o
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr $2 -c $ROOT/dodonaphy_b200/csrc/$f.cu -o $out
  fi
done
nvcc -shared -o $ROOT/_variants/$1.so $ROOT/_variants/$1/prune.o $ROOT/_variants/base_common.o $ROOT/_variants/base_pdm.o $ROOT/_variants/base_nj.o -lcudart
echo $ROOT/_variants/$1.so
