#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python tools/microbench.py vga 4 main > $O/r2c_mb_vga.log 2>&1; cat $O/r2c_mb_vga.log
