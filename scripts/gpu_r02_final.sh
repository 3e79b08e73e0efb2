#!/bin/bash
# last call of round 2: bench lines of the final code + one full capture of the dominant kernel (traffic.json)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/prof
timeout 1200 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; echo "ref rc=$?"
THB_PROFILING=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/prof/final -f python scripts/prof_all.py plan nets eval bwd > gpurun_out/ncu_final.log 2>&1
echo "ncu rc=$?"
python scripts/ncu_raw_summary.py /tmp/prof/final.ncu-rep > gpurun_out/ncu_raw_headline_kernels_r02z3.txt
python scripts/ncu_line_summary.py /tmp/prof/final.ncu-rep k_net_eval 14 > gpurun_out/ncu_lines_k_net_eval_r02z3.txt 2>&1
python scripts/ncu_line_summary.py /tmp/prof/final.ncu-rep k_net_backward 14 > gpurun_out/ncu_lines_k_net_backward_r02z3.txt 2>&1
python scripts/prof_all.py > gpurun_out/times_r02z3.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 3 --inner 2 --no-variants --no-cpu-baseline --no-extras --no-blocks"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 60 --csv --log-file gpurun_out/launches_r02z3.csv $CMD > gpurun_out/ncu_l_r02z3.log 2>&1
echo "launch list rc=$?"
grep -A3 "k_net_eval" gpurun_out/ncu_raw_headline_kernels_r02z3.txt | head -5; cat gpurun_out/times_r02z3.json
