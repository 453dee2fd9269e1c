#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python tools/microbench.py cfg4 6 main p2k1_n1024 p2k1_n512 p2k2_n1024 p2k2_n512 > $O/r2b_mb9_cfg4.log 2>&1; grep -v ckpt $O/r2b_mb9_cfg4.log
