#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
: > $O/r2c_gpu_suite_final.log
for i in 1 2; do echo "== pass $i $(date +%T)" >> $O/r2c_gpu_suite_final.log; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 >> $O/r2c_gpu_suite_final.log; done
cat $O/r2c_gpu_suite_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c_smoke.log 2>&1; tail -1 $O/r2c_smoke.log
