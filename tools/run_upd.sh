python -m pytest tests/test_update_gpu.py tests/test_td3sac_gpu.py -x -q > gpurun_out/t11.log 2>&1; tail -3 gpurun_out/t11.log; grep -n "^E " gpurun_out/t11.log | head -10
python bench.py --workload update --steps 30 > gpurun_out/b11_update.json 2> gpurun_out/b11_update.err; tail -c 300 gpurun_out/b11_update.err
python -c "
import json; d=json.loads(open('gpurun_out/b11_update.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
