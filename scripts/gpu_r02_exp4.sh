#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/exp4.log; : > $L
timeout 600 python -m pytest tests/test_gpu_contraction.py -q -m gpu -x --tb=short >> $L 2>&1; echo "tests rc=$?" >> $L
echo "== staged" >> $L
timeout 600 python scripts/exp_csr.py >> $L 2>&1
echo "== register path backward" >> $L
THB_CSR_BULK_BWD=0 timeout 600 python scripts/exp_csr.py >> $L 2>&1
tail -n 30 $L
