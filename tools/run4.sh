set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
CMD="python bench.py --workload rollout --steps 3 --warmup 3 --n-env 65536"
$CMD > gpurun_out/plain_rollout.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_rollout.csv $CMD > gpurun_out/ncu_l2.log 2>&1
$CMD > gpurun_out/plain_rollout2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:emotion_forward_tc -s 3 -c 1 -o gpurun_out/r01_emotion_tc $CMD > gpurun_out/ncu_f2.log 2>&1
tail -3 gpurun_out/ncu_l2.log gpurun_out/ncu_f2.log
