mkdir -p gpurun_out/la
timeout 900 python -m pytest tests/test_global_gpu.py tests/test_checked_build.py -m gpu -x -q > gpurun_out/la/test2.log 2>&1; echo "tests rc=$?" >> gpurun_out/la/test2.log
timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/la/c5.json 2> gpurun_out/la/c5.err
timeout 300 python bench.py --workload c1 --steps 20 --warmup 3 > gpurun_out/la/c1.json 2> gpurun_out/la/c1.err
bash profiles/scripts/bcr_phase_clocks.sh | tail -1
tail -3 gpurun_out/la/test2.log
python - <<'PY'
import json
for k in ('c5','c1'):
    try:
        d=json.loads(open(f'gpurun_out/la/{k}.json').read().strip().splitlines()[-1])
        pk=d['roofline'].get('per_kernel',{})
        extra=''
        if 'reduced_solve' in pk: extra=' reduced ms/solve %.3f'%(pk['reduced_solve']['ms_per_step']/pk['reduced_solve']['launches_per_step'])
        print(k,'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),extra)
    except Exception as e: print(k,'fail',e)
PY
