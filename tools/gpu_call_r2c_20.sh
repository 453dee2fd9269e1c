#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 > $O/r2c_gpu_suite_rebuilt.log; cat $O/r2c_gpu_suite_rebuilt.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
