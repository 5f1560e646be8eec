python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --workload rollout --steps 10 > gpurun_out/b17_rollout.json 2>gpurun_out/b17.err; tail -c 300 gpurun_out/b17.err
python bench.py --workload td3 --steps 5 > gpurun_out/b17_td3.json 2>>gpurun_out/b17.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01d_launches_td3.csv python bench.py --workload td3 --steps 2 --warmup 3 > gpurun_out/ncu_l11.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01d_launches_esac.csv python bench.py --workload esac --steps 2 --warmup 3 > gpurun_out/ncu_l11b.log 2>&1
python - <<'PY'
import json
for f in ("b17_rollout","b17_td3"):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
PY
