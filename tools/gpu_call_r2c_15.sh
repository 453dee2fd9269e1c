#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_6_host.py tests/test_gpu_2_staging.py -m gpu -x -q 2>&1 | tail -3 > $O/r2c_host_tests.log; cat $O/r2c_host_tests.log
timeout 600 python bench.py --no-others --no-cpu > $O/r2c_bench_d.json 2> $O/r2c_bench_d.err; tail -c 300 $O/r2c_bench_d.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2c_bench_d.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_f32']['ms_per_step'], d['e2e']['loss'], d['e2e_f32']['loss'])
P
timeout 300 python tools/train_step_bench.py gmsd 64 512 10 resident 2>&1 | tail -2
