set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
( TAG=default python profiles/tune.py 1000000; TAG=nosplit LRB_SUM_SPLIT=0 python profiles/tune.py 1000000 ) > gpurun_out/tune.txt 2>&1; cat gpurun_out/tune.txt
