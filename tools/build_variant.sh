#!/bin/bash
# A/B builds of libb200env.so with extra compile-time switches for one source file (kernel tuning experiments).
#   [SRC=uav] tools/build_variant.sh <name> <nvcc flags...>   ->  tools/variants/libb200env_<name>.so  (select with B200ENV_LIB=...)
set -e
name=$1; shift
file=${SRC:-uav}
src=reinforcementlearningplatform_b200/csrc
mkdir -p tools/variants /tmp/var_$name
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
     --expt-relaxed-constexpr -diag-suppress 177 "$@" -c -o /tmp/var_$name/$file.o $src/$file.cu
objs=$(ls $src/*.o | grep -v "\.strict\.o$" | grep -v "/$file\.o\$")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/variants/libb200env_$name.so $objs /tmp/var_$name/$file.o -Xcompiler -fPIC -lcudart
echo built tools/variants/libb200env_$name.so
