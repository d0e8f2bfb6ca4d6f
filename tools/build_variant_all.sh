#!/bin/bash
# A/B build of the WHOLE library with extra compile-time switches (e.g. -DB200_FM64_ESTRIN):
#   tools/build_variant_all.sh <name> <nvcc flags...>  ->  tools/variants/libb200env_<name>.so  (select with B200ENV_LIB=...)
set -e
name=$1; shift
src=reinforcementlearningplatform_b200/csrc
mkdir -p tools/variants /tmp/varall_$name
for f in $src/*.cu; do
  b=$(basename $f .cu); extra=""; [ "$b" = "ugvo" ] && extra="-fmad=false"
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr -diag-suppress 177 $extra "$@" -c -o /tmp/varall_$name/$b.o $f &
done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/variants/libb200env_$name.so /tmp/varall_$name/*.o -Xcompiler -fPIC -lcudart
echo built tools/variants/libb200env_$name.so
