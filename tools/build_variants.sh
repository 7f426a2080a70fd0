#!/bin/bash
# Builds tuning variants of libblu_consensus.so into blutils_b200/variants/ (tile bytes / threads / CTAs per SM).
# usage: tools/build_variants.sh "32768:512:2 20480:320:3 ..."
cd "$(dirname "$0")/../blutils_b200/csrc" || exit 1
mkdir -p ../variants build/var
for v in $1; do
  IFS=: read -r tile thr ctas <<< "$v"
  D="-DBLU_TILE_BYTES=$tile -DBLU_TILE_THREADS=$thr -DBLU_TILE_CTAS=$ctas"
  tag="t${tile}_th${thr}_c${ctas}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $D -Xptxas -v -c blu_kernels.cu -o build/var/k_$tag.o 2> build/var/ptxas_$tag.log || { echo "$tag: kernel build failed"; grep -i "error" build/var/ptxas_$tag.log | head -3; continue; }
  g++ -O2 -std=c++17 -fPIC -I/usr/local/cuda/include $D -c blu_api.cpp -o build/var/a_$tag.o || continue
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libblu_$tag.so build/var/k_$tag.o build/var/a_$tag.o build/blu_taxonomy.o -cudart static -lpthread
  echo "$tag: $(grep -A2 'tile_kernel' build/var/ptxas_$tag.log | grep -o 'Used [0-9]* registers' | head -1), $(grep -A1 'tile_kernel' build/var/ptxas_$tag.log | grep -o '[0-9]* bytes spill stores' | head -1)"
done
