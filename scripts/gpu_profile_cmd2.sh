#!/bin/bash
# full ncu capture of N consecutive kernels matching a regex.  Args: tag kernel-regex skip count command...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
tag=$1; kre=$2; skip=$3; cnt=$4; shift 4
mkdir -p gpurun_out
"$@" > gpurun_out/cmd_plain_$tag.log 2>&1 || { echo "plain command failed"; tail gpurun_out/cmd_plain_$tag.log; exit 1; }
tail -n 3 gpurun_out/cmd_plain_$tag.log
ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c $cnt -o gpurun_out/prof_$tag -f "$@" > gpurun_out/ncu_f_$tag.log 2>&1
echo "full capture rc=$?"; tail -n 3 gpurun_out/ncu_f_$tag.log
