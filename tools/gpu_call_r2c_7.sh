#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
: > $O/r2c_gs_variants.log
for t in gs_cw128 gs_cw32 gs_u8 gs_cw128_u8 gs_nt128; do
  echo "== $t" >> $O/r2c_gs_variants.log
  ADMMTV_LIB=$PWD/admm_deconv_b200/libadmmtv_$t.so timeout 200 python tools/loss_bench.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); g=d['gmsd_loss']; print('gmsd fwd %.4f bwd %.4f' % (g['fwd_train_ms'], g['bwd_ms']))" >> $O/r2c_gs_variants.log
done
cat $O/r2c_gs_variants.log
