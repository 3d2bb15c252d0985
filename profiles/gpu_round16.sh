set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k ont 2>&1 | tail -8
( TAG=ont_flat2 python profiles/tune.py 1000000 ont; TAG=ont_flat1 LRB_SCAN_FLAT=1 python profiles/tune.py 1000000 ont ) > gpurun_out/tune_ont.txt 2>&1; cat gpurun_out/tune_ont.txt
