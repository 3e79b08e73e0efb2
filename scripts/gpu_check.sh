#!/bin/bash
# First-contact GPU run: parity tests (one process per file so a faulting kernel cannot poison the others),
# smoke(), a short bench.  Everything is logged under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc_all=0
for f in tests/test_gpu_contraction.py tests/test_gpu_spans.py tests/test_gpu_basis.py tests/test_gpu_fused.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -x --tb=short > gpurun_out/$n.log 2>&1
  rc=$?; echo "$n rc=$rc"; tail -n 25 gpurun_out/$n.log; [ $rc -ne 0 ] && rc_all=1
done
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -n 15 gpurun_out/bench.err
exit $rc_all
