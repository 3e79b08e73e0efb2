#!/bin/bash
# per-launch kernel durations of a python command: gpu_launch_list_cmd.sh <tag> <skip> <count> <script and args...>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=$1; SKIP=$2; CNT=$3; shift 3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -s $SKIP -c $CNT --csv --log-file gpurun_out/launches_$TAG.csv python "$@" > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$TAG.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]: print(r[ki][:70], r[vi])
PY
