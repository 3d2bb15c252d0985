# Round-2 final visit (one GPU): parity tests, the default bench line, the reference arm, launch lists, full ncu of the heavy kernels.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -3 gpurun_out/r02_bench.err | cut -c1-200
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-400 gpurun_out/r02_bench_reference.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_1m.csv python profiles/run_step.py 1000000 3 > gpurun_out/launch1.log 2>&1
python profiles/launch_table.py gpurun_out/r02_launches_1m.csv > gpurun_out/r02_launch_table_1m.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_10m.csv python profiles/run_step.py 10000000 2 iso 60000 > gpurun_out/launch10.log 2>&1
python profiles/launch_table.py gpurun_out/r02_launches_10m.csv > gpurun_out/r02_launch_table_10m.txt; tail -1 gpurun_out/r02_launch_table_10m.txt
# full capture of the heavy kernels (second step of a 4 M-read data set: loci deep enough for the big-locus fold)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'cigar_scan|classify_row|fold_big|fold_class_rows|fold_prepare' -s 10 -c 12 -o gpurun_out/r02_prof_4m -f python profiles/run_step.py 4000000 2 iso 60000 > gpurun_out/ncu_full4.log 2>&1; tail -2 gpurun_out/ncu_full4.log
