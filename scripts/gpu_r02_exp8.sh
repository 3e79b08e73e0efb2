#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/exp8.log; : > $L
timeout 300 python -m pytest tests/test_gpu_spans.py -q -m gpu -x --tb=short 2>&1 | tail -3 >> $L
for wr in "8 3" "8 0" "4 1" "2 0" "1 0"; do
  for m in serial all after_histogram; do
    SLAB_OVERLAP=$m timeout 300 python scripts/exp_slab.py $wr 2>&1 | tail -1 >> $L
  done
done
cat $L
