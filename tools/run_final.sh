set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bf_default.json 2> gpurun_out/bf_default.err; tail -c 300 gpurun_out/bf_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bf_ref.json 2> gpurun_out/bf_ref.err
python bench.py --workload rollout --steps 10 > gpurun_out/bf_rollout.json 2> gpurun_out/bf_rollout.err
python bench.py --workload update --steps 30 > gpurun_out/bf_update.json 2> gpurun_out/bf_update.err
python bench.py --workload td3 --steps 5 > gpurun_out/bf_td3.json 2> gpurun_out/bf_td3.err
python bench.py --workload sac --steps 5 > gpurun_out/bf_sac.json 2> gpurun_out/bf_sac.err
python bench.py --workload esac --steps 5 > gpurun_out/bf_esac.json 2> gpurun_out/bf_esac.err
python bench.py --workload sim --n-env 65536 --steps 20 --cpu-budget 1 > gpurun_out/bf_sim65536.json 2> gpurun_out/bf_sim65536.err
CMD="python bench.py --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01c_launches_bench_default.csv $CMD > gpurun_out/ncu_lf.log 2>&1
python - <<'PY'
import json
for f in ("bf_default","bf_ref","bf_rollout","bf_update","bf_td3","bf_sac","bf_esac","bf_sim65536"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('roofline') or {}).get('frac'))
        for o in d.get('other_workloads',[]): print('   ', o['metric'], o['value'], o['ms_per_step'], o['roofline'].get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
