set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k ont 2>&1 | tail -4
( TAG=ont_flat2_nosmem python profiles/tune.py 1000000 ont ) > gpurun_out/tune_ont.txt 2>&1; cat gpurun_out/tune_ont.txt
