#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 600 python bench.py > $O/r2c_bench_c.json 2> $O/r2c_bench_c.err; tail -c 300 $O/r2c_bench_c.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2c_bench_c.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e'], d.get('e2e_f32'))
P
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c_bench_ref_c.json 2> $O/r2c_bench_ref_c.err; head -c 200 $O/r2c_bench_ref_c.json
