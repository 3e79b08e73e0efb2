#!/bin/bash
# ncu evidence for the dominant kernel: launch list (durations) + one full capture. Args: tag
cd "${GRAFT_REPO_ROOT:-/root/repo}"
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline"
$CMD > gpurun_out/prof_plain_$tag.json 2> gpurun_out/prof_plain_$tag.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 60 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_l_$tag.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_net_eval -s 3 -c 1 -o gpurun_out/prof_$tag -f $CMD > gpurun_out/ncu_f_$tag.log 2>&1
echo "full capture rc=$?"; tail -n 5 gpurun_out/ncu_f_$tag.log
