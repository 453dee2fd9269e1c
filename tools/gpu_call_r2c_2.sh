#!/bin/bash
# 2-CTA-cluster dim-2 pass for N = 2048: forward parity suite, then the variants on configs[3]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_0_forward.py -m gpu -x -q 2>&1 | tail -15 > $O/r2c_fwd_tests.log; cat $O/r2c_fwd_tests.log
timeout 600 python tools/microbench.py cfg4 4 noclu main clu4_16_mb2 clu4_16_u1 clu4_8 clu8_16 clu2_8 > $O/r2c_mb_cfg4.log 2>&1; grep -v ckpt-fwd $O/r2c_mb_cfg4.log
