set -x
python bench.py --workload sim --steps 20 --cpu-budget 1 > gpurun_out/b9_sim.json 2> gpurun_out/b9_sim.err; tail -c 300 gpurun_out/b9_sim.err
CMD="python bench.py --steps 2 --warmup 3 --cpu-budget 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01d_launches_bench_default.csv $CMD > gpurun_out/ncu_l9.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_sim_warp -s 3 -c 1 -o gpurun_out/r01d_env_sim_warp_4096 -f python bench.py --workload sim --steps 2 --warmup 3 --cpu-budget 1 > gpurun_out/ncu_f9a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"env_sim_warp|env_finish" -s 6 -c 2 -o gpurun_out/r01d_env_sim_warp_65536 -f python bench.py --workload sim --n-env 65536 --steps 2 --warmup 3 --cpu-budget 1 > gpurun_out/ncu_f9b.log 2>&1
ls -la gpurun_out/*.ncu-rep
