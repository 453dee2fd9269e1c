#!/bin/bash
# round-2b call 1: suite, reversed dim-1 order and 2048 variants, ncu --set full of the 2048 kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2b_gpu_tests.log 2>&1; tail -2 $O/r2b_gpu_tests.log
python tools/microbench.py cfg4 6 main rev1 tr11_4_mb2 pf2_148 n11_256 > $O/r2b_mb_cfg4.log 2>&1; grep -v ckpt $O/r2b_mb_cfg4.log
python tools/microbench.py cfg2 10 main rev1 main rev1 > $O/r2b_mb_cfg2.log 2>&1; cat $O/r2b_mb_cfg2.log
python tools/microbench.py cfg3 10 main rev1 > $O/r2b_mb_cfg3.log 2>&1; cat $O/r2b_mb_cfg3.log
python profiles/ncu_target.py cfg4 4 fwd > $O/r2b_plain_cfg4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_dim1_fwd|k_dim2" -s 9 -c 4 -f -o $O/r2b_cfg4 python profiles/ncu_target.py cfg4 4 fwd > $O/r2b_ncu_cfg4.log 2>&1
ADMMTV_LIB=$PWD/admm_deconv_b200/libadmmtv_tr11_4_mb2.so ncu --set full --clock-control none --import-source on -k regex:"k_dim2" -s 5 -c 1 -f -o $O/r2b_cfg4_tr4 python profiles/ncu_target.py cfg4 4 fwd > $O/r2b_ncu_cfg4_tr4.log 2>&1
ls -la $O/*.ncu-rep
