#!/bin/bash
# Side-by-side builds of libunimat.so with -D knobs on tprf.cu only:  tools/build_variants.sh NAME="-DFOO -DBAR" ...
# -> uni_b200/variants/libunimat_NAME.so (git-ignored; select with UNIMAT_LIB)
set -e
cd "$(dirname "$0")/../uni_b200/csrc"
mkdir -p ../variants build
OBJS=$(ls build/*.o | grep -vx "build/tprf\.o" | grep -v "build/tprf_v_")
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr $flags -c tprf.cu -o build/tprf_v_$name.o 2> build/tprf_v_$name.log \
    && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libunimat_$name.so $OBJS build/tprf_v_$name.o -lnccl -lcuda && echo "built $name" ) &
done
wait
