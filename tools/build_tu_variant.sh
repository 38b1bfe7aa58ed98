#!/bin/bash
# Fast A/B build of ONE translation unit: tools/build_tu_variant.sh NAME TU [-DFLAG ...]
#   recompiles numbacs_b200/csrc/TU.cu with the flags and links it against the product build's
#   other objects -> build/variants/libb200cs_NAME.so
set -e
name=$1; tu=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/build/variants/obj_fast; mkdir -p "$obj"
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    "$@" -c "$root/numbacs_b200/csrc/$tu.cu" -o "$obj/${tu}_$name.o"
others=$(ls "$root"/numbacs_b200/csrc/build/*.o | grep -v "/$tu.o")
nvcc -shared -o "$root/build/variants/libb200cs_$name.so" "$obj/${tu}_$name.o" $others -gencode arch=compute_100a,code=sm_100a
echo "$root/build/variants/libb200cs_$name.so"
