#!/bin/bash
# one full ncu capture of one kernel of a script: gpu_r02_ncu_one.sh <kernel regex> <tag> <python script> [skip]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/prof
K=$1; TAG=$2; SCRIPT=$3; SKIP=${4:-2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o /tmp/prof/$TAG -f python $SCRIPT > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_$TAG.log
python scripts/ncu_raw_summary.py /tmp/prof/$TAG.ncu-rep > gpurun_out/ncu_raw_$TAG.txt
python scripts/ncu_line_summary.py /tmp/prof/$TAG.ncu-rep $K 24 > gpurun_out/ncu_lines_$TAG.txt 2>&1
cat gpurun_out/ncu_raw_$TAG.txt gpurun_out/ncu_lines_$TAG.txt
ncu -i /tmp/prof/$TAG.ncu-rep --page source --csv --kernel-name regex:$K > /tmp/prof/$TAG.src.csv 2>/dev/null
python scripts/ncu_src_summary.py /tmp/prof/$TAG.src.csv > gpurun_out/ncu_sass_mix_$TAG.txt 2>&1
cat gpurun_out/ncu_sass_mix_$TAG.txt
