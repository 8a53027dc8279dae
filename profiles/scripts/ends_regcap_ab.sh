#!/bin/bash
# ends_regcap_ab.sh: line_endpoints_kernel with 80 registers (shipped), 64 (-DLINE_EP_MIN_BLOCKS=8) and 48 (=10)
cd "$(dirname "$0")/../.."
for v in "" build/ab/ep8.so build/ab/ep10.so; do
  if [ -n "$v" ]; then export RSPL_BA_LIB=$v; else unset RSPL_BA_LIB; fi
  timeout 200 python bench.py --workload ends --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lib', '${v:-shipped}', 'kernel ms', round(d['ms_per_step'],4), 'parity', d.get('parity_check'))"
done
