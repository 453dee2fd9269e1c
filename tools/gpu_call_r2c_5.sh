#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r2c_gpu_suite.log; cat $O/r2c_gpu_suite.log
timeout 300 python tools/loss_bench.py > $O/r2c_loss_bench.log 2>&1; tail -5 $O/r2c_loss_bench.log
