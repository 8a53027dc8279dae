#!/bin/bash
# frame_ab.sh LIB...: device time of the K7 workloads for A/B builds of the library (build_variant.sh), via bench.py
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lib in "$@"; do
  for w in c2p c2 frame1; do
    RSPL_BA_LIB=$PWD/$lib python bench.py --workload $w --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
d=d.get('workloads',{}).get('$w',d) if '$w'!='c2' else d
print('$lib','$w','ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['ms_per_step'],4),'fp',d.get('result_fingerprint'))"
  done
done
