set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py > gpurun_out/b13_default.json 2> gpurun_out/b13_default.err; tail -c 300 gpurun_out/b13_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b13_ref.json 2> gpurun_out/b13_ref.err; tail -c 200 gpurun_out/b13_ref.err
python bench.py --workload td3 --steps 5 > gpurun_out/b13_td3.json 2> gpurun_out/b13_td3.err; tail -c 300 gpurun_out/b13_td3.err
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv --log-file gpurun_out/r01c_launches_td3_warm.csv python bench.py --workload td3 --steps 2 --warmup 3 > gpurun_out/ncu_l13.log 2>&1
python - <<'PY'
import json
for f in ("b13_default","b13_ref","b13_td3"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('roofline') or {}).get('frac'))
        for o in d.get('other_workloads',[]): print('   ', o['metric'], o['value'], o['ms_per_step'], o['roofline'].get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
