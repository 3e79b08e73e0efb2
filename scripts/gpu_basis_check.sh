#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for ppt in 1 2; do
  export THB_BASIS_PPT=$ppt
  timeout 900 python -m pytest tests/test_gpu_basis.py tests/test_gpu_large.py -q -m gpu -x --tb=short 2>&1 | tail -n 3
  python bench.py --steps 5 --warmup 3 --no-variants --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('PPT=$ppt', json.dumps(d['extras']['THB_basis_fns']))"
done
