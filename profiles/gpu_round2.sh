# Round-2 profile visit: launch lists (configs[1] and a deep 10 M-read data set), full ncu of the heavy kernels, DRAM traffic per kernel.
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_1m.csv python profiles/run_step.py 1000000 3 > gpurun_out/launch1.log 2>&1
python profiles/launch_table.py gpurun_out/r02_launches_1m.csv > gpurun_out/r02_launch_table_1m.txt; cat gpurun_out/r02_launch_table_1m.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_10m.csv python profiles/run_step.py 10000000 2 iso 60000 > gpurun_out/launch10.log 2>&1
python profiles/launch_table.py gpurun_out/r02_launches_10m.csv > gpurun_out/r02_launch_table_10m.txt; cat gpurun_out/r02_launch_table_10m.txt
# full capture of the heavy kernels of the second step of the deep data set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cigar_scan|classify_row|fold_big|fold_class_rows|fold_prepare|fold_relrep|sum_phase1|sum_exon_insert|compact_gather' -s 40 -c 14 -o gpurun_out/r02_prof_10m -f python profiles/run_step.py 10000000 2 iso 60000 > gpurun_out/ncu_full10.log 2>&1; tail -3 gpurun_out/ncu_full10.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cigar_stream' -s 1 -c 1 -o gpurun_out/r02_prof_ont -f python profiles/run_step.py 2000000 2 ont 60000 > gpurun_out/ncu_full_ont.log 2>&1; tail -3 gpurun_out/ncu_full_ont.log
