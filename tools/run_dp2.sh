set -x
TR="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 python -m pytest tests/test_peer_gpu.py -x -q 2>&1 | tail -3
$TR bench.py --gpus 2 --workload update --steps 30 > gpurun_out/b12_update_n2_peer.json 2> gpurun_out/b12_update_n2_peer.err; tail -c 300 gpurun_out/b12_update_n2_peer.err
$TR bench.py --gpus 2 --workload update --steps 30 --exchange nccl > gpurun_out/b12_update_n2_nccl.json 2> gpurun_out/b12_update_n2_nccl.err; tail -c 300 gpurun_out/b12_update_n2_nccl.err
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b12_default_n2.json 2> gpurun_out/b12_default_n2.err; tail -c 300 gpurun_out/b12_default_n2.err
python - <<'PY'
import json
for f in ("b12_update_n2_peer","b12_update_n2_nccl","b12_default_n2"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d.get('n_gpus'), d['value'], d['ms_per_step'], d['e2e']['value'])
        for o in d.get('other_workloads',[]): print('   ', o['metric'], o['value'], o['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
