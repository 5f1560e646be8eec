set -x
python -m pytest tests/test_env_step_gpu.py -x -q -m gpu 2>&1 | tail -15
for m in warp thread; do
  python bench.py --n-env 4096 --steps 20 --sim-mode $m --cpu-budget 1 > gpurun_out/b_pid4096_$m.json 2>gpurun_out/b_pid4096_$m.err
  python bench.py --n-env 65536 --actions random --steps 10 --sim-mode $m --cpu-budget 1 > gpurun_out/b_rand65536_$m.json 2>gpurun_out/b_rand65536_$m.err
  python bench.py --n-env 262144 --steps 10 --sim-mode $m --cpu-budget 1 > gpurun_out/b_pid262144_$m.json 2>gpurun_out/b_pid262144_$m.err
done
tail -c 400 gpurun_out/*.err
