#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 python -m pytest tests/test_gpu_0_forward.py tests/test_gpu_6_host.py -m gpu -x -q 2>&1 | tail -1
