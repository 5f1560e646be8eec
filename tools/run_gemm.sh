python -m pytest tests/test_update_gpu.py tests/test_td3sac_gpu.py tests/test_esac_gpu.py -x -q 2>&1 | tail -2
for w in update td3 esac; do python bench.py --workload $w --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['ms_per_step'])"; done
