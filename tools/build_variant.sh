#!/bin/bash
# A/B build: tools/build_variant.sh NAME "-DFLAG ..." file.cu [file2.cu ...]  ->  volpick_b200/libvolpick_b200_NAME.so
# (the named sources recompiled with the extra flags, every other object taken from the stock build; use with VP_LIB_PATH)
set -eu
name=$1; flags=$2; shift 2
python -m volpick_b200.build > /dev/null
objs=""
for o in volpick_b200/_obj/*.o; do
  b=$(basename $o .o); skip=0
  for f in "$@"; do [ "$(basename $f .cu)" = "$b" ] && skip=1; done
  [ $skip = 0 ] && objs="$objs $o"
done
for f in "$@"; do
  mkdir -p volpick_b200/_obj/variant; o=volpick_b200/_obj/variant/$(basename $f .cu)_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $flags -c $f -o $o
  objs="$objs $o"
done
nvcc -shared -o volpick_b200/libvolpick_b200_$name.so $objs -cudart static -gencode arch=compute_100a,code=sm_100a
echo volpick_b200/libvolpick_b200_$name.so
