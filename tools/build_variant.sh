#!/bin/bash
# builds a kernel variant for A/B runs: tools/build_variant.sh <name> <sed expression over csrc/*.cuh> -> slam.net_b200/_build/variants/lib_<name>.so
# (select it with CS_B200_LIB=<path>)
name=$1; shift
src=/tmp/cs_variant_$name
rm -rf $src; mkdir -p $src; cp -r slam.net_b200/csrc/* $src/
sed -i "s#\.\./\.\./include/coreslam_b200.h#$PWD/include/coreslam_b200.h#" $src/*.cu $src/*.cuh $src/*.h $src/host/* 2>/dev/null
for e in "$@"; do sed -i "$e" $src/*.cuh $src/*.cu; done
mkdir -p slam.net_b200/_build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared -Iinclude -Xptxas -v \
  -o slam.net_b200/_build/variants/lib_$name.so $src/cs_api.cu > /tmp/cs_variant_$name.log 2>&1; grep -A2 cs_wedge_kernelILb1 /tmp/cs_variant_$name.log | grep "spill\|Used"; grep -i "error" /tmp/cs_variant_$name.log | head -3
ls -la slam.net_b200/_build/variants/lib_$name.so | awk '{print $5, $9}'
