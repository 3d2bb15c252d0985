set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
( TAG=default python profiles/tune.py 1000000; TAG=crt256 LRB_CR_THREADS=256 python profiles/tune.py 1000000; TAG=crt64 LRB_CR_THREADS=64 python profiles/tune.py 1000000 ) > gpurun_out/tune.txt 2>&1; cat gpurun_out/tune.txt
timeout 900 python profiles/cli_e2e.py 600000 > gpurun_out/cli_e2e.json 2> gpurun_out/cli_e2e.err; cat gpurun_out/cli_e2e.json; tail -3 gpurun_out/cli_e2e.err
