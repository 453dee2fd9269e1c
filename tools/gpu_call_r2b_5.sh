#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2b_gpu_tests5.log 2>&1; tail -5 $O/r2b_gpu_tests5.log
timeout 120 python tools/microbench.py cfg2 10 main notma1 --iso > $O/r2b_mb_cfg2_iso_tma1.log 2>&1; grep -v ckpt $O/r2b_mb_cfg2_iso_tma1.log
timeout 120 python tools/microbench.py cfg5 10 main notma1 > $O/r2b_mb_cfg5_tma1.log 2>&1; cat $O/r2b_mb_cfg5_tma1.log
