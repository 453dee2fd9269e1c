#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
: > $O/r2c_gpu_suite_x6.log
for i in 1 2 3 4 5 6; do echo "== pass $i $(date +%T)" >> $O/r2c_gpu_suite_x6.log; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 >> $O/r2c_gpu_suite_x6.log; done
cat $O/r2c_gpu_suite_x6.log
