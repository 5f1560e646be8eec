set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/b8_default.json 2> gpurun_out/b8_default.err; tail -c 300 gpurun_out/b8_default.err
python bench.py --workload rollout --steps 10 > gpurun_out/b8_rollout.json 2> gpurun_out/b8_rollout.err; tail -c 300 gpurun_out/b8_rollout.err
python bench.py --workload update --steps 20 > gpurun_out/b8_update.json 2> gpurun_out/b8_update.err; tail -c 300 gpurun_out/b8_update.err
CMD="python bench.py --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01b_launches_bench_default.csv $CMD > gpurun_out/ncu_l8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:emotion_forward_tc -s 3 -c 1 -o gpurun_out/r01b_emotion_tc -f python tools/bench_emotion.py 65536 > gpurun_out/ncu_f8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_sim_warp -s 3 -c 1 -o gpurun_out/r01b_env_sim_warp -f python bench.py --workload sim --steps 2 --warmup 3 > gpurun_out/ncu_f8b.log 2>&1
python - <<'PY'
import json
for f in ("b8_default","b8_rollout","b8_update"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('frac'))
        for o in d.get('other_workloads',[]): print('   ', o['metric'], o['value'], o['ms_per_step'], o['roofline'].get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
