#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_1_losses.py -m gpu -x -q 2>&1 | tail -4 > $O/r2c_loss_tests.log; cat $O/r2c_loss_tests.log
timeout 300 python tools/loss_bench.py > $O/r2c_loss_bench_b.log 2>&1; tail -2 $O/r2c_loss_bench_b.log
