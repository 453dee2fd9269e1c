#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_0_forward.py -m gpu -x -q 2>&1 | tail -5 > $O/r2c_fwd_tests.log; cat $O/r2c_fwd_tests.log
timeout 600 python tools/microbench.py w4096 4 noclu main > $O/r2c_mb_w4096.log 2>&1; grep -v ckpt-fwd $O/r2c_mb_w4096.log
timeout 600 python -m pytest tests/test_gpu_1_losses.py -m gpu -x -q 2>&1 | tail -5 > $O/r2c_loss_tests.log; cat $O/r2c_loss_tests.log
