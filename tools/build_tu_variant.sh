#!/bin/bash
# Fast A/B build of ONE OR MORE translation units: tools/build_tu_variant.sh NAME TU[,TU...] [-DFLAG ...]
#   recompiles numbacs_b200/csrc/TU.cu with the flags and links them against the product build's
#   other objects -> build/variants/libb200cs_NAME.so
set -e
name=$1; tus=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/build/variants/obj_fast; mkdir -p "$obj"
others=$(ls "$root"/numbacs_b200/csrc/build/*.o)
mine=""
for tu in ${tus//,/ }; do
    nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
        "$@" -c "$root/numbacs_b200/csrc/$tu.cu" -o "$obj/${tu}_$name.o" &
    others=$(echo "$others" | grep -v "/$tu.o")
    mine="$mine $obj/${tu}_$name.o"
done
wait
nvcc -shared -o "$root/build/variants/libb200cs_$name.so" $mine $others -gencode arch=compute_100a,code=sm_100a
echo "$root/build/variants/libb200cs_$name.so"
