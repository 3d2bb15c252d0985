set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
( TAG=default python profiles/tune.py 1000000; TAG=minb10 LRB_CR_MINB=10 python profiles/tune.py 1000000 ) > gpurun_out/tune.txt 2>&1; cat gpurun_out/tune.txt
