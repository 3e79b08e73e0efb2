#!/bin/bash
# every GPU test file in its own process + the driver's own invocation at the end
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rc_all=0
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py); timeout 900 python -m pytest $f -q -m gpu -x --tb=short > gpurun_out/$n.log 2>&1; rc=$?
  echo "$n rc=$rc"; tail -n 3 gpurun_out/$n.log; [ $rc -ne 0 ] && { rc_all=1; tail -n 40 gpurun_out/$n.log; }
done
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "all-in-one rc=$?"; tail -n 3 gpurun_out/pytest_gpu_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
exit $rc_all
