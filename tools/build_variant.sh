#!/bin/bash
# Builds an instrumented copy of libsmplb.so: tools/build_variant.sh <name> <extra nvcc flags...>
# -> tools/micro/ab/libsmplb_<name>.so (git-ignored, travels to the GPU box); use with SMPLB_LIB=<path>.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/human-pose-estimation_b200/csrc
out=$root/tools/micro/ab/$name
mkdir -p $out
for f in $src/*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$root/include -I$src "$@" -c $f -o $out/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/tools/micro/ab/libsmplb_$name.so $out/*.o -lcudart -ldl
echo built $root/tools/micro/ab/libsmplb_$name.so
