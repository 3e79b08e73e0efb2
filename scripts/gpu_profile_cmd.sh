#!/bin/bash
# one full ncu capture of a kernel of an arbitrary command.  Args: tag kernel-regex skip command...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
tag=$1; kre=$2; skip=$3; shift 3
mkdir -p gpurun_out
"$@" > gpurun_out/cmd_plain_$tag.log 2>&1 || { echo "plain command failed"; tail gpurun_out/cmd_plain_$tag.log; exit 1; }
tail -n 3 gpurun_out/cmd_plain_$tag.log
ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -o gpurun_out/prof_$tag -f "$@" > gpurun_out/ncu_f_$tag.log 2>&1
echo "full capture rc=$?"; tail -n 3 gpurun_out/ncu_f_$tag.log
