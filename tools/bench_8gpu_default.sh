set -x
N=${1:-8}
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 3 --cpu-budget 2 > gpurun_out/b18_default_n$N.json 2> gpurun_out/b18_default_n$N.err; tail -c 300 gpurun_out/b18_default_n$N.err
$TR bench.py --gpus $N --workload sim --n-env 65536 --steps 20 --cpu-budget 1 > gpurun_out/b18_sim65536_n$N.json 2> gpurun_out/b18_sim65536_n$N.err; tail -c 300 gpurun_out/b18_sim65536_n$N.err
python - <<PY
import json
for f in ("b18_default_n$N","b18_sim65536_n$N"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('n_gpus'), d['value'], d['ms_per_step'], d['e2e']['value'])
        for o in d.get('other_workloads',[]): print('   ', o['metric'], o['value'], o['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
