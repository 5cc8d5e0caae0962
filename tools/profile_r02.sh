#!/bin/bash
# ncu evidence of round 2 (run under gpurun from the repo root): --set full captures of the three tcgen05 kernels and the
# launch lists of one step / one call.  Every profiled command is run plain first (exit code checked) as the recipe asks.
set -u
O=gpurun_out
python tools/one_step.py > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 6 -c 1 -f -o $O/r02q_tower_full python tools/one_step.py > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
python tools/one_call.py 1 > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lat_tower_kernel -s 3 -c 1 -f -o $O/r02q_lat_full python tools/one_call.py 1 > $O/ncu2.log 2>&1; tail -1 $O/ncu2.log
python tools/one_step.py 2048 fp32 3 > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 45 -c 1 -f -o $O/r02q_split_full python tools/one_step.py 2048 fp32 3 > $O/ncu3.log 2>&1; tail -1 $O/ncu3.log
python tools/one_step.py 2048 bf16 3 > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 5 --csv --log-file $O/r02q_launches_2048.csv python tools/one_step.py 2048 bf16 3 > $O/ncu4.log 2>&1
python tools/one_step.py 2048 fp32 2 > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 130 --csv --log-file $O/r02q_launches_2048_fp32.csv python tools/one_step.py 2048 fp32 2 > $O/ncu5.log 2>&1
python tools/one_call.py 1 > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 10 --csv --log-file $O/r02q_launches_n1.csv python tools/one_call.py 1 > $O/ncu6.log 2>&1
ls -la $O/*.ncu-rep
