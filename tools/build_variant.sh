#!/bin/bash
# Builds a variant of the library with extra -D flags for ONE translation unit (default gen_vertices.cu):
#   tools/build_variant.sh <name> "<extra nvcc flags>" [file.cu]   ->  build/variants/lib_<name>.so   (use with BT_LIB_PATH)
set -e
name=$1; extra=$2; unit=${3:-gen_vertices.cu}
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/bevy_terrain_b200/csrc
out=$root/build/variants
mkdir -p $out
obj=$out/${unit%.cu}_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v $extra -c $src/$unit -o $obj 2> $out/${unit%.cu}_$name.ptxas.log
others=$(ls $src/*.o | grep -v "/${unit%.cu}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/lib_$name.so $obj $others -lz -ldl
grep -A1 "gen_vertices_kernelILb0ELi16E\|gen_vertices_kernelILb0ELi22E" $out/${unit%.cu}_$name.ptxas.log | grep "registers\|spill" || true
echo "built $out/lib_$name.so"
