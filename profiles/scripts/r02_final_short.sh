mkdir -p gpurun_out/final2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final2/test.log 2>&1; echo "tests rc=$?" >> gpurun_out/final2/test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final2/smoke.log
timeout 900 python bench.py > gpurun_out/final2/n1.json 2> gpurun_out/final2/n1.err; echo "bench rc=$?" >> gpurun_out/final2/test.log
tail -3 gpurun_out/final2/test.log; tail -1 gpurun_out/final2/smoke.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final2/n1.json').read().strip().splitlines()[-1])
def line(k,v): print(k,'ms',round(v['ms_per_step'],3),'e2e',round(v['e2e']['ms_per_step'],3),'roof',round(v['roofline']['frac'],3), 'shim', v.get('shim_latency',{}).get('median_us'))
line('c2',d)
for k,v in d.get('workloads',{}).items(): line(k,v)
PY
