#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "every_fft_length or any_size or tma_pipelined" > $O/r2b_gpu_tests3.log 2>&1; tail -2 $O/r2b_gpu_tests3.log
python tools/microbench.py cfg4 6 main rb1 rb2_rev1 rb2_pfn rb2_cpre main > $O/r2b_mb_cfg4_rb.log 2>&1; grep -v ckpt $O/r2b_mb_cfg4_rb.log
