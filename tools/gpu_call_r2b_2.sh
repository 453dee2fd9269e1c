#!/bin/bash
# round-2b call 2: dim-2 tile swizzle (TR < 16)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q -k "every_fft_length or any_size or tma_pipelined" > $O/r2b_gpu_tests2.log 2>&1; tail -2 $O/r2b_gpu_tests2.log
python tools/microbench.py cfg4 6 main noswz2 tr11_4_mb2 tr11_4_mb3 main > $O/r2b_mb_cfg4_swz.log 2>&1; grep -v ckpt $O/r2b_mb_cfg4_swz.log
