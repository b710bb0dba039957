#!/bin/bash
# Build the native library of another git revision into storm_b200/lib_<name>.so for same-box A/B runs:
#   tools/build_variant.sh HEAD base ; STORM_B200_LIB=storm_b200/lib_base.so python bench.py ...
set -e
rev=$1; name=$2
tmp=$(mktemp -d)
git archive "$rev" storm_b200/csrc include | tar -x -C "$tmp"
objs=""
for f in engine kernels_small conv_simt conv_tc attn_tc; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177 \
    -c "$tmp/storm_b200/csrc/$f.cu" -o "$tmp/$f.o" &
  objs="$objs $tmp/$f.o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o "storm_b200/lib_$name.so" $objs -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf "$tmp"
echo "storm_b200/lib_$name.so"
