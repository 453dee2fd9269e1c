#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "every_fft_length or any_size or tma_pipelined or vs_oracle" > $O/r2b_gpu_tests4.log 2>&1; tail -5 $O/r2b_gpu_tests4.log
timeout 120 python tools/microbench.py cfg2 10 main notma1 main notma1 > $O/r2b_mb_cfg2_tma1.log 2>&1; grep -v ckpt $O/r2b_mb_cfg2_tma1.log
timeout 120 python tools/microbench.py cfg4 6 main notma1 > $O/r2b_mb_cfg4_tma1.log 2>&1; grep -v ckpt $O/r2b_mb_cfg4_tma1.log
timeout 120 python tools/microbench.py cfg3 10 main notma1 > $O/r2b_mb_cfg3_tma1.log 2>&1; grep -v ckpt $O/r2b_mb_cfg3_tma1.log
