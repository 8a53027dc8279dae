#!/bin/bash
# dense_timing.sh: time per reduced solve of the dense Cholesky (dense_chol.cuh) against the cyclic reduction on the
# same banded systems (RSPL_BA_DENSE_FULL=1 forces the dense path), via bench.py's profiled pass
cd "$(dirname "$0")/../.."
for cfg in "160 16000 1600" "600 300000 30000" "2000 1000000 100000"; do
  set -- $cfg
  for full in 0 1; do
    if [ $full = 1 ]; then export RSPL_BA_DENSE_FULL=1; else unset RSPL_BA_DENSE_FULL; fi
    python bench.py --workload c5 --kf $1 --points $2 --lines $3 --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
pk=d['roofline']['per_kernel']
rs=pk['reduced_solve']
print('kf',$1,'dense_full',$full,'ms/step',round(d['ms_per_step'],2),'reduced_solve ms/solve',round(rs['ms_per_step']/rs['launches_per_step'],3),'assemble',round(pk.get('dense_assemble',{}).get('ms_per_step',0),2))"
  done
done
