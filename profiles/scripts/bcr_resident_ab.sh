#!/bin/bash
# bcr_resident_ab.sh: global-BA tests + C5 with the cyclic-reduction variants (RSPL_BA_BCR_SLABS=1: slab-staged update kernel)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/bcr
timeout 600 python -m pytest tests/test_global_gpu.py tests/test_checked_build.py -m gpu -x -q > gpurun_out/bcr/test.log 2>&1; echo "tests rc=$?" >> gpurun_out/bcr/test.log
for v in 0 1; do
  if [ $v = 1 ]; then export RSPL_BA_BCR_SLABS=1; else unset RSPL_BA_BCR_SLABS; fi
  timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/bcr/c5_slabs$v.json 2> gpurun_out/bcr/c5_slabs$v.err
done
unset RSPL_BA_BCR_SLABS
tail -3 gpurun_out/bcr/test.log
python - <<'PY'
import json
for v in (0,1):
    try:
        d=json.loads(open(f'gpurun_out/bcr/c5_slabs{v}.json').read().strip().splitlines()[-1])
        rs=d['roofline']['per_kernel']['reduced_solve']
        print('slabs',v,'ms/step',round(d['ms_per_step'],2),'reduced ms/solve',round(rs['ms_per_step']/rs['launches_per_step'],3))
    except Exception as e: print(v,'fail',e)
PY
