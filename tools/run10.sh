for m in 0 1 2; do
  ERLFILL_STEP_HOST_MODE=$m python bench.py --workload sim --steps 50 --warmup 3 --cpu-budget 0.5 > gpurun_out/b11_$m.json 2>gpurun_out/b11_$m.err
  python -c "
import json; d=json.loads(open('gpurun_out/b11_$m.json').read().strip().splitlines()[-1]); print('mode $m', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
ERLFILL_STEP_HOST_MODE=2 python -m pytest tests/test_env_step_gpu.py -x -q -k "graph_replay" 2>&1 | tail -2
