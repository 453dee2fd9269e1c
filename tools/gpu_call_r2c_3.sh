#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_0_forward.py -m gpu -x -q 2>&1 | tail -5 > $O/r2c_fwd_tests.log; cat $O/r2c_fwd_tests.log
timeout 600 python tools/microbench.py cfg4 4 main > $O/r2c_mb_cfg4b.log 2>&1; grep -v ckpt-fwd $O/r2c_mb_cfg4b.log
timeout 200 python profiles/ncu_target.py cfg4 4 fwd > $O/r2c_plain_cfg4.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dim2c" -s 3 -c 1 -f -o $O/r2c_cfg4_cluster python profiles/ncu_target.py cfg4 4 fwd > $O/r2c_ncu_cfg4.log 2>&1
tail -3 $O/r2c_ncu_cfg4.log
