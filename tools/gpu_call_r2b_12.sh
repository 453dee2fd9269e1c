#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
: > $O/r2b_gpu_suite_x10.log
for i in 1 2 3 4 5 6 7 8 9 10; do echo "== pass $i $(date +%T)" >> $O/r2b_gpu_suite_x10.log; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 >> $O/r2b_gpu_suite_x10.log; done
cat $O/r2b_gpu_suite_x10.log
timeout 300 python tools/graph_bench.py > $O/r2b_graph_bench.log 2>&1; tail -8 $O/r2b_graph_bench.log
timeout 200 python profiles/ncu_target.py cfg4 4 fwd > $O/r2b_plain_cfg4b.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dim1_fwd|k_dim2" -s 9 -c 4 -f -o $O/r2b_cfg4_after python profiles/ncu_target.py cfg4 4 fwd > $O/r2b_ncu_cfg4b.log 2>&1
