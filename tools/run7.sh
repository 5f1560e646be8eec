set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b7_default_n2.json 2> gpurun_out/b7_default_n2.err; tail -c 300 gpurun_out/b7_default_n2.err
$TR bench.py --gpus 2 --workload update --steps 20 > gpurun_out/b7_update_n2.json 2> gpurun_out/b7_update_n2.err; tail -c 300 gpurun_out/b7_update_n2.err
$TR bench.py --gpus 2 --workload td3 --steps 5 > gpurun_out/b7_td3_n2.json 2> gpurun_out/b7_td3_n2.err; tail -c 300 gpurun_out/b7_td3_n2.err
$TR bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/b7_ref_n2.json 2> gpurun_out/b7_ref_n2.err; tail -c 300 gpurun_out/b7_ref_n2.err
python - <<'PY'
import json
for f in ("b7_default_n2","b7_update_n2","b7_td3_n2","b7_ref_n2"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('n_gpus'), d['value'], d['ms_per_step'], d['e2e']['value'])
        for o in d.get('other_workloads',[]): print('   ', o['metric'], o['value'], o['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
