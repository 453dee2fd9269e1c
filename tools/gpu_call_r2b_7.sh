#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 200 python tools/microbench.py cfg2 10 main rev1 main rev1 > $O/r2b_rev_cfg2.log 2>&1; grep -v ckpt $O/r2b_rev_cfg2.log
timeout 200 python tools/microbench.py cfg3 10 main rev1 > $O/r2b_rev_cfg3.log 2>&1; grep -v ckpt $O/r2b_rev_cfg3.log
timeout 200 python tools/microbench.py cfg4 6 main rev1 > $O/r2b_rev_cfg4.log 2>&1; grep -v ckpt $O/r2b_rev_cfg4.log
timeout 600 python bench.py > $O/r2b_bench_a.json 2> $O/r2b_bench_a.err; tail -c 600 $O/r2b_bench_a.err; head -c 1500 $O/r2b_bench_a.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-others --no-cpu > $O/r2b_plain_bench.log 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none -c 2000 --csv --log-file $O/r2b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-others --no-cpu > $O/r2b_ncu_bench.log 2>&1
timeout 200 python profiles/ncu_target.py cfg2 6 train > $O/r2b_plain_target.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dim1_bwd|k_dim1_fwd|k_dim2" -s 30 -c 8 -f -o $O/r2b_prof_cfg2 python profiles/ncu_target.py cfg2 6 train > $O/r2b_ncu_full.log 2>&1
ls -la $O/*.ncu-rep
