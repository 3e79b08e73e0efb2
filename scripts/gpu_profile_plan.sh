#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
tag=${1:-plan}
CMD="python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_plan -s 10 -c 5 -o gpurun_out/prof_$tag -f $CMD > gpurun_out/ncu_p_$tag.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/ncu_p_$tag.log
