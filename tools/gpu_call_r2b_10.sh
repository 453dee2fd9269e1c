#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2b_gpu_tests10.log 2>&1; tail -3 $O/r2b_gpu_tests10.log
timeout 200 python tools/microbench.py cfg3 10 main main > $O/r2b_mb10_cfg3.log 2>&1; cat $O/r2b_mb10_cfg3.log
timeout 200 python tools/microbench.py cfg3_share 10 main > $O/r2b_mb10_cfg3s.log 2>&1; cat $O/r2b_mb10_cfg3s.log
