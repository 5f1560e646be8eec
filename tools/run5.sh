set -x
python bench.py > gpurun_out/b5_default.json 2> gpurun_out/b5_default.err; tail -c 300 gpurun_out/b5_default.err
CMD="python bench.py --workload update --steps 3 --warmup 3"
$CMD > gpurun_out/plain_update.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_update.csv $CMD > gpurun_out/ncu_l3.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/b5_default.json'))
print(d['value'], d['e2e'], d.get('issue_roofline',{}).get('frac'))
for o in d['other_workloads']: print(o['metric'], o['value'], o['ms_per_step'], o['roofline']['frac'])
PY
