#!/bin/bash
# 8-GPU record on the final library: bench line with the all-reduce inside the step (own arm), GPU suite's distributed tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2c_bench_n8.json 2> $O/r2c_bench_n8.err; tail -c 300 $O/r2c_bench_n8.err; head -c 500 $O/r2c_bench_n8.json
timeout 300 python -m pytest tests/test_gpu_5_dist.py -m gpu -x -q 2>&1 | tail -2 > $O/r2c_dist_tests_n8.log; cat $O/r2c_dist_tests_n8.log
