set -x
python bench.py --workload rollout --steps 10 > gpurun_out/b3_rollout.json 2> gpurun_out/b3_rollout.err
python bench.py --workload rollout --steps 10 --n-env 262144 > gpurun_out/b3_rollout262144.json 2> gpurun_out/b3_rollout262144.err
python bench.py --workload sim --n-env 65536 --actions random --steps 10 --cpu-budget 1 > gpurun_out/b3_simrand65536.json 2> gpurun_out/b3_simrand.err
tail -c 600 gpurun_out/b3_*.err
cat gpurun_out/b3_rollout.json gpurun_out/b3_rollout262144.json
