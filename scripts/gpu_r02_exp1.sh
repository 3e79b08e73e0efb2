#!/bin/bash
# round 2, session 3: vector histogram kernel (parity + timing), nets beside the scatter pass, e2e chunk / ramp sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=gpurun_out/exp1.log; : > $L
timeout 600 python -m pytest tests/test_gpu_spans.py tests/test_gpu_baseline_configs.py tests/test_gpu_large.py -q -m gpu -x --tb=short >> $L 2>&1; echo "tests rc=$?" >> $L
echo "== exp_step" >> $L
THB_COUNT_VEC=0 timeout 300 python scripts/exp_step.py >> $L 2>&1
timeout 300 python scripts/exp_step.py >> $L 2>&1
for nc in 3 4 6; do THB_OVERLAP_NETS_CTAS=$nc timeout 300 python scripts/exp_step.py >> $L 2>&1; done
for sc in 4 6 8; do THB_SCATTER_CTAS_PER_SM=$sc timeout 300 python scripts/exp_step.py >> $L 2>&1; done
THB_SCATTER_CTAS_PER_SM=6 THB_OVERLAP_NETS_CTAS=4 timeout 300 python scripts/exp_step.py >> $L 2>&1
THB_SCATTER_CTAS_PER_SM=4 THB_OVERLAP_NETS_CTAS=5 timeout 300 python scripts/exp_step.py >> $L 2>&1
echo "== e2e" >> $L
CFGS='"8 0" "16 0" "16 3" "24 4" "32 0" "32 4"' 
for cfg in "8 0" "16 0" "16 3" "24 4" "32 0" "32 4"; do
  set -- $cfg
  python bench.py --steps 10 --warmup 3 --no-variants --no-extras --no-cpu-baseline --no-blocks --e2e-chunks $1 --e2e-ramp $2 2>/dev/null > /tmp/e2e.json
  python -c "import json; d=json.load(open('/tmp/e2e.json')); print('chunks', $1, 'ramp', $2, 'e2e ms', round(d['e2e']['ms_per_step'],3), 'duplex copies alone', round(d['e2e']['host_copies_alone']['duplex_ms'],3), d['e2e']['check_max_abs_diff_vs_device_path'], 'pass ms', d['config']['ms_per_pass'])" >> $L 2>&1
done
tail -n 40 $L
