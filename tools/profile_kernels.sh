ncu --set full --clock-control none --import-source on -k regex:replay_insert_ring -s 12 -c 1 -o gpurun_out/r01d_replay_insert -f python bench.py --workload replay --steps 2 --warmup 3 > gpurun_out/ncu_f13a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:replay_gather -s 14 -c 1 -o gpurun_out/r01d_replay_gather -f python bench.py --workload replay --steps 2 --warmup 3 > gpurun_out/ncu_f13b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 130 -c 3 -o gpurun_out/r01d_td3_gemm -f python bench.py --workload td3 --steps 2 --warmup 3 > gpurun_out/ncu_f13c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:actor_forward_kernel -s 4 -c 1 -o gpurun_out/r01d_actor_forward -f python bench.py --workload rollout --steps 2 --warmup 3 > gpurun_out/ncu_f13d.log 2>&1
ls -la gpurun_out/*.ncu-rep
