# One GPU visit: parity tests, bench line, per-stage A/B (env knobs), launch list, full ncu of the heavy kernels.
# Usage: bash profiles/gpu_round.sh [quick]
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
( TAG=default python profiles/tune.py 1000000; for v in $TUNE_VARIANTS; do env TAG=$v $v python profiles/tune.py 1000000; done ) > gpurun_out/tune.txt 2>&1; cat gpurun_out/tune.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python profiles/run_step.py 1000000 3 > gpurun_out/launch.log 2>&1; tail -2 gpurun_out/launch.log
python profiles/launch_table.py gpurun_out/launches.csv > gpurun_out/launch_table.txt; cat gpurun_out/launch_table.txt
[ "$1" = quick ] && exit 0
# full ncu capture of the heavy kernels of the last step (skip the first two steps)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cigar_scan|classify_row|fold_relrep|sum_exon|sum_bed|fold_seq|fold_prepare|sum_phase|compact_gather' -s 30 -c 15 -o gpurun_out/prof -f python profiles/run_step.py 1000000 3 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
