N=${1:-8}
TR="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
$TR bench.py --gpus $N --workload td3 --steps 5 > gpurun_out/b29_td3_n$N.json 2> gpurun_out/b29_td3_n$N.err; tail -c 300 gpurun_out/b29_td3_n$N.err
$TR bench.py --gpus $N --workload sac --steps 5 > gpurun_out/b29_sac_n$N.json 2> gpurun_out/b29_sac_n$N.err; tail -c 300 gpurun_out/b29_sac_n$N.err
python - <<PY
import json
for f in ("b29_td3_n$N","b29_sac_n$N"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('n_gpus'), d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
