#!/bin/bash
# final profiles of round 2: ncu full capture of every kernel of the path (summarised on the box: the report itself is
# too large to travel), launch list of bench steps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/prof
THB_PROFILING=1 timeout 1800 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/prof/prof_r02z2 -f python scripts/prof_all.py > gpurun_out/ncu_r02z2.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_r02z2.log; ls -la /tmp/prof
python scripts/ncu_raw_summary.py /tmp/prof/prof_r02z2.ncu-rep > gpurun_out/ncu_raw_all_kernels_r02z2.txt
for k in k_net_eval k_cell_nets k_plan_count_vec k_plan_scatter "k_csr_forward<" k_csr_forward_bulk k_csr_backward_bulk k_basis_rows k_net_backward k_cell_unfold k_grid_eval; do
  python scripts/ncu_line_summary.py /tmp/prof/prof_r02z2.ncu-rep $k 14 > gpurun_out/ncu_lines_${k}_r02z2.txt 2>&1
done
python scripts/ncu_src_summary.py /tmp/prof/prof_r02z2.ncu-rep k_net_eval > gpurun_out/ncu_sass_mix_net_eval_r02z2.txt 2>&1
ncu -i /tmp/prof/prof_r02z2.ncu-rep --page raw --csv --kernel-name regex:k_net_eval > gpurun_out/ncu_raw_csv_net_eval_r02z2.csv 2>/dev/null
CMD="python bench.py --steps 2 --warmup 3 --inner 2 --no-variants --no-cpu-baseline --no-extras --no-blocks"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 60 --csv --log-file gpurun_out/launches_r02z2.csv $CMD > gpurun_out/ncu_l_r02z2.log 2>&1
echo "launch list rc=$?"; du -sh gpurun_out
