#!/bin/bash
# frame_sweep.sh LIB N...: device time of the point-only K7 batch for several batch sizes (occupancy / tail study)
cd "$(dirname "$0")/../.."
lib=$1; shift
for n in "$@"; do
  RSPL_BA_LIB=$PWD/$lib python bench.py --workload c2p --frames $n --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
d=d.get('workloads',{}).get('c2p',d)
print('$lib', 'frames', $n, 'ms', round(d['ms_per_step'],4), 'us/frame', round(1e3*d['ms_per_step']/$n,4))"
done
