set -x
python bench.py --steps 20 --cpu-budget 3 > gpurun_out/b2_sim.json 2> gpurun_out/b2_sim.err
python bench.py --workload rollout --steps 5 > gpurun_out/b2_rollout.json 2> gpurun_out/b2_rollout.err
python bench.py --workload rollout --n-env 4096 --steps 5 > gpurun_out/b2_rollout4096.json 2> gpurun_out/b2_rollout4096.err
python bench.py --workload update --steps 20 > gpurun_out/b2_update.json 2> gpurun_out/b2_update.err
python bench.py --workload update --batch 128 --replay 100000 --steps 20 > gpurun_out/b2_update128.json 2> gpurun_out/b2_update128.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b2_ref.json 2> gpurun_out/b2_ref.err
tail -c 300 gpurun_out/b2_*.err
