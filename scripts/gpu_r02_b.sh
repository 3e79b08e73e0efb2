#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for f in tests/test_gpu_fit.py tests/test_gpu_c5.py; do
  n=$(basename $f .py); timeout 900 python -m pytest $f -q -m gpu --tb=short > gpurun_out/$n.log 2>&1; rc=$?
  echo "$n rc=$rc"; tail -n 3 gpurun_out/$n.log; [ $rc -ne 0 ] && tail -n 60 gpurun_out/$n.log
done
timeout 600 python scripts/prof_all.py > gpurun_out/times_r02a.json 2> gpurun_out/times_r02a.err; echo "times rc=$?"; cat gpurun_out/times_r02a.json; tail -n 5 gpurun_out/times_r02a.err
THB_PROFILING=1 timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r02a -f python scripts/prof_all.py > gpurun_out/ncu_r02a.log 2>&1
echo "ncu rc=$?"; tail -n 5 gpurun_out/ncu_r02a.log; ls -la gpurun_out/prof_r02a.ncu-rep
