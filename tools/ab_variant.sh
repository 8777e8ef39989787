#!/bin/bash
# Build the in-tree sources with extra nvcc flags into tools/ab/<name>.so (development: A/B of compile-time knobs).
#   tools/ab_variant.sh <name> -DWT_CUT_TAIL=0 ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root"
python -m swgl_b200.build >/dev/null
mkdir -p tools/ab /tmp/swgl_var_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
     -Xcompiler -fPIC -I include -I swgl_b200/csrc "$@" -c swgl_b200/csrc/swgl_dev.cu -o /tmp/swgl_var_$name/dev.o 2>/dev/null
g++ -shared -o tools/ab/$name.so /tmp/swgl_var_$name/dev.o swgl_b200/_build/swgl_host.c.o swgl_b200/_build/swgl_glsl.c.o swgl_b200/_build/swgl_jit.cpp.o \
    -Wl,-Bsymbolic -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -lcudart -ldl -lpthread -lstdc++
echo tools/ab/$name.so
