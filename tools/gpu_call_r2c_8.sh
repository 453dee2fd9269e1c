#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:"k_gmsd|k_ssim" -c 60 --csv --log-file $O/r2c_launches_losses.csv python tools/loss_bench.py > $O/r2c_ncu_losses.log 2>&1
tail -2 $O/r2c_ncu_losses.log
