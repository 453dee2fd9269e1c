#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_6_host.py -m gpu -x -q 2>&1 | tail -3 > $O/r2c_host_tests.log; cat $O/r2c_host_tests.log
