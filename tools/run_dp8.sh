set -x
N=${1:-8}
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 170 python -m pytest tests/test_peer_gpu.py -x -q 2>&1 | tail -3
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/b16_default_n$N.json 2> gpurun_out/b16_default_n$N.err; tail -c 300 gpurun_out/b16_default_n$N.err
$TR bench.py --gpus $N --workload update --steps 30 --exchange nccl > gpurun_out/b16_update_n${N}_nccl.json 2> gpurun_out/b16_update_n${N}_nccl.err; tail -c 300 gpurun_out/b16_update_n${N}_nccl.err
python - <<PY
import json
for f in ("b16_default_n$N","b16_update_n${N}_nccl"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('n_gpus'), d['value'], d['ms_per_step'], d['e2e']['value'])
        for o in d.get('other_workloads',[]): print('   ', o['metric'], o['value'], o['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
