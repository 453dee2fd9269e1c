#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python tools/microbench.py cfg2 10 main t18p1 c4mb3 t6p1 t10_mb5 main > $O/r2b_sweep_cfg2.log 2>&1; cat $O/r2b_sweep_cfg2.log
timeout 200 python tools/microbench.py cfg3 10 main t18c8 t34c8 mb8_4 > $O/r2b_sweep_cfg3.log 2>&1; cat $O/r2b_sweep_cfg3.log
timeout 200 python tools/microbench.py cfg4 6 main n11_512_t10 > $O/r2b_sweep_cfg4.log 2>&1; cat $O/r2b_sweep_cfg4.log
