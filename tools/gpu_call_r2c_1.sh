#!/bin/bash
# 2-GPU sanity after the TMA dim-1 change: the distributed GPU tests and the bench line with the all-reduce in the step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_5_dist.py -m gpu -x -q 2>&1 | tail -3 > $O/r2c_dist_tests_n2.log; cat $O/r2c_dist_tests_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2c_bench_n2.json 2> $O/r2c_bench_n2.err; tail -c 400 $O/r2c_bench_n2.err; head -c 600 $O/r2c_bench_n2.json
