#!/bin/bash
# Fast A/B build: recompile only flowmap_dg.cu with the given -D flags and link it against the
# product build's other objects (numbacs_b200/csrc/build/*.o, which must be current).
#   tools/build_dg_variant.sh NAME [-DFLAG ...]  ->  build/variants/libb200cs_NAME.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/build/variants/obj_fast; mkdir -p "$obj"
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    "$@" -c "$root/numbacs_b200/csrc/flowmap_dg.cu" -o "$obj/flowmap_dg_$name.o"
others=$(ls "$root"/numbacs_b200/csrc/build/*.o | grep -v "/flowmap_dg.o")
nvcc -shared -o "$root/build/variants/libb200cs_$name.so" "$obj/flowmap_dg_$name.o" $others -gencode arch=compute_100a,code=sm_100a
echo "$root/build/variants/libb200cs_$name.so"
