set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python profiles/run_step.py 1000000 3 > gpurun_out/launch.log 2>&1; tail -2 gpurun_out/launch.log
python profiles/launch_table.py gpurun_out/launches.csv > gpurun_out/launch_table.txt; cat gpurun_out/launch_table.txt
