#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_6_host.py tests/test_gpu_1_losses.py -m gpu -x -q 2>&1 | tail -3 > $O/r2c_host_tests.log; cat $O/r2c_host_tests.log
timeout 600 python bench.py --no-others --no-cpu > $O/r2c_bench_e.json 2> $O/r2c_bench_e.err; tail -c 300 $O/r2c_bench_e.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2c_bench_e.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_f32']['ms_per_step'], d['e2e']['loss'], d['e2e_f32']['loss'])
P
