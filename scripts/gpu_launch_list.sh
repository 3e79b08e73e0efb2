#!/bin/bash
# per-launch durations of the library kernels of one bench step (cold cache, serialised).  Args: tag
cd "${GRAFT_REPO_ROOT:-/root/repo}"
tag=${1:-ll}
CMD="python bench.py --steps 2 --warmup 3 --no-variants --no-cpu-baseline --no-extras"
$CMD > gpurun_out/ll_plain_$tag.json 2> gpurun_out/ll_plain_$tag.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 40 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_l_$tag.log 2>&1
echo "launch list rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$tag.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-14:-7]: print(r[4][:60].ljust(60), r[-1])
PY
