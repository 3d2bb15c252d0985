set -x
mkdir -p gpurun_out
( TAG=default python profiles/tune.py 1000000; TAG=crsort LRB_CR_SORT=1 python profiles/tune.py 1000000 ) > gpurun_out/tune.txt 2>&1; cat gpurun_out/tune.txt
LRB_CR_SORT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
LRB_IO_TRACE=1 timeout 600 python profiles/cli_e2e.py 300000 > gpurun_out/cli_e2e.json 2> gpurun_out/cli_e2e.err; cat gpurun_out/cli_e2e.json; grep "lrb io" gpurun_out/cli_e2e.err | tail -40
