#!/bin/bash
# tuning build: only the (4,4,4) tile kernels, extra -D flags; output thb_diff_b200/_tune/lib_<tag>.so
set -e
tag=$1; shift
cd /root/repo; mkdir -p thb_diff_b200/_tune/$tag
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Iinclude -Ithb_diff_b200/csrc -DTHB_ONLY_444 $@"
for f in thb_util thb_eval_csr thb_eval_bulk thb_grid thb_spans thb_dense thb_fit thb_tile thb_tile_inst_444; do
  nvcc $F -c thb_diff_b200/csrc/$f.cu -o thb_diff_b200/_tune/$tag/$f.o &
done
wait
nvcc -shared -o thb_diff_b200/_tune/lib_$tag.so thb_diff_b200/_tune/$tag/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built $tag
