#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for v in pf1 pf3 pf4; do echo $v; THB_LIB=thb_diff_b200/_tune/lib_$v.so timeout 300 python scripts/prof_all.py csr32 csr64 2>&1 | tail -1; done
