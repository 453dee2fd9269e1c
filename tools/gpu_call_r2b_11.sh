#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2b_smoke.log 2>&1; tail -3 $O/r2b_smoke.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py > $O/r2b_memcheck.log 2>&1; tail -4 $O/r2b_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py > $O/r2b_racecheck.log 2>&1; tail -4 $O/r2b_racecheck.log
timeout 600 python bench.py > $O/r2b_bench_b.json 2> $O/r2b_bench_b.err; tail -c 300 $O/r2b_bench_b.err; head -c 400 $O/r2b_bench_b.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-others --no-cpu > $O/r2b_plain_bench.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none -c 2000 --csv --log-file $O/r2b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-others --no-cpu > $O/r2b_ncu_bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2b_bench_ref.json 2> $O/r2b_bench_ref.err; head -c 600 $O/r2b_bench_ref.json
