#!/bin/bash
# ends_check.sh: parity tests of the line endpoint refresh + its bench line (both arms) + an ncu capture of the kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/ends
timeout 600 python -m pytest tests/test_maplines.py tests/test_triangulate.py -m gpu -x -q > gpurun_out/ends/test.log 2>&1; echo "tests rc=$?" >> gpurun_out/ends/test.log
timeout 300 python bench.py --workload ends --steps 5 --warmup 3 > gpurun_out/ends/ends.json 2> gpurun_out/ends/ends.err
timeout 300 python bench.py --workload tri --steps 5 --warmup 3 > gpurun_out/ends/tri.json 2>> gpurun_out/ends/ends.err
timeout 300 python bench.py --impl reference --workload ends --steps 3 --warmup 1 > gpurun_out/ends/ends_ref.json 2>> gpurun_out/ends/ends.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:line_endpoints_kernel -c 1 -o gpurun_out/ends/ncu_ends python bench.py --workload ends --steps 1 --warmup 3 > gpurun_out/ends/ncu.log 2>&1
ncu -i gpurun_out/ends/ncu_ends.ncu-rep --page raw --csv > gpurun_out/ends/ncu_ends_raw.csv 2>/dev/null
tail -4 gpurun_out/ends/test.log
python - <<'PY'
import json
for k in ("ends", "tri"):
    d = json.loads(open(f"gpurun_out/ends/{k}.json").read().strip().splitlines()[-1])
    print(k, "kernel ms", round(d["ms_per_step"], 4), "roof", round(d["roofline"]["frac"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "parity", d.get("parity_check"), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
