#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/exp7.log; : > $L
python scripts/exp_slab.py 8 3 >> $L 2>&1
python scripts/exp_slab.py 8 0 >> $L 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|Memset|memset" -s 600 -c 24 --csv --log-file gpurun_out/launches_slab8_3.csv python scripts/exp_slab.py 8 3 > /dev/null 2>&1
python - <<'PY' >> $L
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_slab8_3.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]: print(r[ki][:60], r[vi])
PY
cat $L
