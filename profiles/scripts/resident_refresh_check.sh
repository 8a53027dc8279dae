mkdir -p gpurun_out/res
timeout 900 python -m pytest tests/test_maplines.py tests/test_checked_build.py -m gpu -x -q > gpurun_out/res/test.log 2>&1; echo "tests rc=$?" >> gpurun_out/res/test.log
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/res/c4.json 2> gpurun_out/res/c4.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/res/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/res/smoke.log
tail -5 gpurun_out/res/test.log; tail -3 gpurun_out/res/smoke.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/res/c4.json').read().strip().splitlines()[-1])
print('c4 ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
print(d.get('post_ba_line_refresh'))
PY
