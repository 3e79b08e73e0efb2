#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ls thb_diff_b200/_tune/ 2>&1 | head
timeout 600 python -m pytest tests/test_gpu_contraction.py -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/prof_all.py csr32 csr64 2>&1 | tail -1
for v in noatom perm; do echo $v; THB_LIB=thb_diff_b200/_tune/lib_$v.so timeout 300 python scripts/prof_all.py csr32 csr64 2>&1 | tail -1; done
for v in m12 m20 m24; do echo $v; THB_LIB=thb_diff_b200/_tune/lib_$v.so timeout 300 python scripts/prof_all.py eval 2>&1 | tail -1; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-variants --no-blocks --no-cpu-baseline > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_r02i.json; tail -5 gpurun_out/bench_r02i.err
