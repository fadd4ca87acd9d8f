#!/bin/bash
# A/B sweep of the streamed tensor-core kernel knobs on a B200 (run under gpurun)
set -u
OPS="pw48f pw24f dhtf dhts dhta"
echo "== default"; python tools/profile_ops.py $OPS 3
for pf in 0 24 48 192; do echo "== HNO_TC_PREFETCH_KB=$pf"; HNO_TC_PREFETCH_KB=$pf python tools/profile_ops.py $OPS 3; done
for pr in 0 2; do echo "== HNO_TC_L2PROMO=$pr"; HNO_TC_L2PROMO=$pr python tools/profile_ops.py $OPS 3; done
echo "== promo 2 + prefetch 0"; HNO_TC_L2PROMO=2 HNO_TC_PREFETCH_KB=0 python tools/profile_ops.py $OPS 3
