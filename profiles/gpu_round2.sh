set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
( TAG=default python profiles/tune.py 1000000; TAG=fused LRB_FOLD_FUSED=1 python profiles/tune.py 1000000 ) > gpurun_out/tune.txt 2>&1; cat gpurun_out/tune.txt
timeout 600 python profiles/cli_e2e.py 300000 > gpurun_out/cli_e2e.json 2> gpurun_out/cli_e2e.err; cat gpurun_out/cli_e2e.json; tail -3 gpurun_out/cli_e2e.err
nproc
