set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
( TAG=ont_flat python profiles/tune.py 1000000 ont; TAG=ont_warp LRB_SCAN_FLAT=0 python profiles/tune.py 1000000 ont ) > gpurun_out/tune_ont.txt 2>&1; cat gpurun_out/tune_ont.txt
