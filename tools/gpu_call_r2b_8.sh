#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2b_gpu_tests8.log 2>&1; tail -3 $O/r2b_gpu_tests8.log
timeout 200 python tools/microbench.py cfg2 10 main main > $O/r2b_mb8_cfg2.log 2>&1; cat $O/r2b_mb8_cfg2.log
timeout 200 python tools/microbench.py cfg3 10 main > $O/r2b_mb8_cfg3.log 2>&1; cat $O/r2b_mb8_cfg3.log
timeout 200 python tools/microbench.py cfg4 6 main > $O/r2b_mb8_cfg4.log 2>&1; cat $O/r2b_mb8_cfg4.log
