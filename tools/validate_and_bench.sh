python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bf2_default.json 2> gpurun_out/bf2_default.err; tail -c 300 gpurun_out/bf2_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bf2_ref.json 2> gpurun_out/bf2_ref.err
python bench.py --workload rollout --steps 10 > gpurun_out/bf2_rollout.json 2> gpurun_out/bf2_rollout.err
python bench.py --workload update --steps 30 > gpurun_out/bf2_update.json 2> gpurun_out/bf2_update.err
python bench.py --workload sim --n-env 65536 --steps 20 --cpu-budget 1 > gpurun_out/bf2_sim65536.json 2> gpurun_out/bf2_sim65536.err
python bench.py --workload replay --steps 20 > gpurun_out/bf2_replay.json 2> gpurun_out/bf2_replay.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01e_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --cpu-budget 1 > gpurun_out/ncu_lf2.log 2>&1
python - <<'PY'
import json
for f in ("bf2_default","bf2_ref","bf2_rollout","bf2_update","bf2_sim65536","bf2_replay"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('roofline') or {}).get('frac'))
        for o in d.get('other_workloads',[]): print('   ', o['metric'], o['value'], o['ms_per_step'], o['roofline'].get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
