#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for f in tests/test_gpu_spans.py tests/test_gpu_basis.py tests/test_gpu_fused.py; do
  n=$(basename $f .py); timeout 900 python -m pytest $f -q -m gpu -x --tb=short > gpurun_out/$n.log 2>&1; echo "$n rc=$?"; tail -n 4 gpurun_out/$n.log
done
for ppt in 2 1; do
  THB_FUSED_PPT=$ppt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ppt$ppt.json 2> gpurun_out/bench_ppt$ppt.err
  echo "ppt=$ppt rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_ppt$ppt.json"))
r=d["roofline"]; print("value %.3e e2e %.3e step %.3f ms kernel %.3f plan %.3f frac %.3f frac8d %.3f peak %.1f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["kernel_ms"],r["plan_ms"],r["frac"],r["frac_vs_dense_model_8d"],r["peak"]))
print(d.get("variants"))
PY
done
