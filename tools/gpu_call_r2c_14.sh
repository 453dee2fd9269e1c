#!/bin/bash
# final 1-GPU evidence: GPU suite x2, smoke, ncu launch list of the bench command (after it exited 0 without ncu)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
: > $O/r2c_gpu_suite_final.log
for i in 1 2; do echo "== pass $i $(date +%T)" >> $O/r2c_gpu_suite_final.log; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 >> $O/r2c_gpu_suite_final.log; done
cat $O/r2c_gpu_suite_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c_smoke.log 2>&1; tail -2 $O/r2c_smoke.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-others --no-cpu > $O/r2c_plain_bench.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none -c 2000 --csv --log-file $O/r2c_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-others --no-cpu > $O/r2c_ncu_bench.log 2>&1
tail -2 $O/r2c_ncu_bench.log | cut -c1-200
