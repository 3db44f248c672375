#!/bin/bash
# round-2 baseline of the shipped kernel: gpu tests, bench line, ncu launch list, one full capture (+ spill / tail metrics)
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_gputests.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
EXTRA=smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max,sm__cycles_elapsed.avg,smsp__cycles_active.avg
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2a_ncu_l.log 2>&1
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:machine_kernel -s 4 -c 1 -o gpurun_out/prof_r2a_bench -f python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_ncu_f.log 2>&1
ls -la gpurun_out | tail -5
