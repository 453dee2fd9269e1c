#!/bin/bash
# full GPU suite x3 on the round-2c library, smoke, bench (own arm + reference arm), ncu launch list of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
: > $O/r2c_gpu_suite_x3.log
for i in 1 2 3; do echo "== pass $i $(date +%T)" >> $O/r2c_gpu_suite_x3.log; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 >> $O/r2c_gpu_suite_x3.log; done
cat $O/r2c_gpu_suite_x3.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c_smoke.log 2>&1; tail -3 $O/r2c_smoke.log
timeout 600 python bench.py > $O/r2c_bench.json 2> $O/r2c_bench.err; tail -c 300 $O/r2c_bench.err; head -c 300 $O/r2c_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c_bench_ref.json 2> $O/r2c_bench_ref.err; head -c 300 $O/r2c_bench_ref.json
