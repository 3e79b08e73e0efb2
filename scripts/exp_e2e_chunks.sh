#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for cfg in ${CFGS:-"8 0" "10 2" "12 3" "14 4" "12 4"}; do
  set -- $cfg
  python bench.py --steps 10 --warmup 3 --no-variants --no-extras --no-cpu-baseline --e2e-chunks $1 --e2e-ramp $2 2>/dev/null > /tmp/e2e.json
  python -c "import json; d=json.load(open('/tmp/e2e.json')); print('chunks', $1, 'ramp', $2, 'e2e ms', round(d['e2e']['ms_per_step'],3), 'duplex copies alone', round(d['e2e']['host_copies_alone']['duplex_ms'],3), d['e2e']['check_max_abs_diff_vs_device_path'])"
done
