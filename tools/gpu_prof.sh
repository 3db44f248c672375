#!/bin/bash
# usage: bash tools/gpu_prof.sh TAG [bench args...]   -> gpurun_out/TAG_bench.json, TAG_launches.csv, prof_TAG.ncu-rep
TAG=$1; shift
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || { tail -5 gpurun_out/${TAG}_bench.err; exit 1; }
EXTRA=smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max,sm__cycles_elapsed.avg,smsp__cycles_active.avg
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu "$@" > gpurun_out/${TAG}_ncu_l.log 2>&1
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:machine_kernel -s 4 -c 1 -o gpurun_out/prof_${TAG} -f python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/${TAG}_ncu_f.log 2>&1
cat gpurun_out/${TAG}_bench.json | head -c 600
