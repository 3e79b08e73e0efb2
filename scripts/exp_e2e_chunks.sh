#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for c in ${CHUNKS:-8 16 24 32}; do
  python bench.py --steps 10 --warmup 3 --no-variants --no-extras --no-cpu-baseline --e2e-chunks $c 2>/dev/null > /tmp/e2e_$c.json
  python -c "import json; d=json.load(open('/tmp/e2e_$c.json')); print('chunks', $c, 'e2e ms', round(d['e2e']['ms_per_step'],3), 'duplex copies alone', round(d['e2e']['host_copies_alone']['duplex_ms'],3))"
done
