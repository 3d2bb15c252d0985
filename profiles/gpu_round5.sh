set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
( TAG=default python profiles/tune.py 1000000; TAG=relw0 LRB_FOLD_RELW=0 python profiles/tune.py 1000000; TAG=slots24 LRB_FOLD_SLOTS=24 python profiles/tune.py 1000000; TAG=slots16 LRB_FOLD_SLOTS=16 python profiles/tune.py 1000000 ) > gpurun_out/tune.txt 2>&1; cat gpurun_out/tune.txt
timeout 600 python profiles/cli_e2e.py 300000 > gpurun_out/cli_e2e.json 2> gpurun_out/cli_e2e.err; cat gpurun_out/cli_e2e.json; tail -3 gpurun_out/cli_e2e.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
