#!/bin/bash
# r02_final.sh: the closing pass of round 2 on one B200 -- full GPU test suite, smoke, both bench arms (stored under
# profiles/bench_r02/), then ncu: launch list of C5, `--set full` captures of the cyclic-reduction kernels, the line
# endpoint kernel and the triangulation kernel, and the Schur-class traffic of C5 (profiles/traffic_c5.json).
cd "$(dirname "$0")/../.."
O=gpurun_out/final
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/test.log 2>&1; echo "tests rc=$?" >> $O/test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py --impl reference > $O/ref_n1.json 2> $O/ref_n1.err; echo "ref rc=$?" >> $O/test.log
timeout 900 python bench.py > $O/n1.json 2> $O/n1.err; echo "bench rc=$?" >> $O/test.log
NCU="ncu --clock-control none"
summ() { python profiles/ncu_summary.py $O/$1.ncu-rep > $O/$1.txt; }
$NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/r02_launches_c5_final2.csv python bench.py --workload c5 --steps 1 --warmup 1 > $O/ncu_c5.log 2>&1
$NCU --set full --import-source on -k regex:"bcr_eliminate|bcr_update_resident|bcr_backsub|bcr_root" -s 24 -c 4 -o $O/r02_bcr3 python profiles/scripts/r02_c5_target.py 2000 > $O/ncu_b.log 2>&1
summ r02_bcr3; rm -f $O/r02_bcr3.ncu-rep
$NCU --set full --import-source on -k regex:"line_endpoints_kernel" -c 1 -o $O/r02_ends python bench.py --workload ends --steps 1 --warmup 3 > $O/ncu_e.log 2>&1
summ r02_ends; rm -f $O/r02_ends.ncu-rep
$NCU --set full --import-source on -k regex:"triangulate_points_kernel" -c 1 -o $O/r02_tri python bench.py --workload tri --steps 1 --warmup 3 > $O/ncu_t.log 2>&1
summ r02_tri; rm -f $O/r02_tri.ncu-rep
bash profiles/scripts/r02_profile_c5.sh > $O/profile_c5.log 2>&1
cp gpurun_out/r02_c5_schur.txt $O/ 2>/dev/null
tail -4 $O/test.log; tail -2 $O/smoke.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final/n1.json').read().strip().splitlines()[-1])
def line(k,v): print(k,'ms',round(v['ms_per_step'],3),'e2e',round(v['e2e']['ms_per_step'],3),'roof',round(v['roofline']['frac'],3), 'parity',v.get('parity_check'))
line('c2',d)
for k,v in d.get('workloads',{}).items(): line(k,v)
PY
