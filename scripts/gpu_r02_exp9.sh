#!/bin/bash
# points per work item: 128 / 256 (default) / 512 / 1024 on the headline step and the fit step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/exp9.log; : > $L
for ppi in 128 256 512 1024; do
  echo "== ppi $ppi" >> $L
  THB_POINTS_PER_ITEM=$ppi timeout 300 python scripts/exp_step.py 2>&1 | tail -1 | cut -c100-330 >> $L
  THB_POINTS_PER_ITEM=$ppi timeout 300 python scripts/prof_all.py eval bwd fit 2>&1 | tail -1 >> $L
done
cat $L
