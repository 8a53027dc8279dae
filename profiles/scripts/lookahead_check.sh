#!/bin/bash
# lookahead_check.sh: full GPU suite + the latency / solve-bound workloads after a change to the shared CTA Cholesky
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/la
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/la/test.log 2>&1; echo "tests rc=$?" >> gpurun_out/la/test.log
for k in c1 c3 c4 c5; do
  st=20; [ $k = c4 ] && st=3; [ $k = c5 ] && st=3
  timeout 400 python bench.py --workload $k --steps $st --warmup 3 > gpurun_out/la/$k.json 2> gpurun_out/la/$k.err
done
tail -3 gpurun_out/la/test.log
python - <<'PY'
import json
for k in ('c1','c3','c4','c5'):
    try:
        d=json.loads(open(f'gpurun_out/la/{k}.json').read().strip().splitlines()[-1])
        pk=d['roofline'].get('per_kernel',{})
        extra=''
        if 'reduced_solve' in pk: extra=' reduced_solve class ms %.3f'%(pk['reduced_solve']['ms_per_step'])
        print(k,'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),extra)
    except Exception as e: print(k,'fail',e)
PY
