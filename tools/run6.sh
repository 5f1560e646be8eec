set -x
python bench.py --workload td3 --steps 5 > gpurun_out/b6_td3.json 2> gpurun_out/b6_td3.err; tail -c 400 gpurun_out/b6_td3.err
python bench.py --workload sac --steps 5 > gpurun_out/b6_sac.json 2> gpurun_out/b6_sac.err; tail -c 400 gpurun_out/b6_sac.err
python - <<'PY'
import json
for f in ("b6_td3","b6_sac"):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['share_of_step'])
    except Exception as e: print(f, 'ERR', e)
PY
