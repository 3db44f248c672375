#!/bin/bash
# usage: bash tools/gpu_ext_prof.sh TAG   -> config 3 / config 4 bench lines, launch lists, one ncu capture of the cloud kernel and of
# the longest spectral path wave (gpurun_out/TAG_*)
TAG=$1
mkdir -p gpurun_out
for c in config3 config4; do
  python bench.py --workload $c --no-cpu --steps 3 > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err || tail -3 gpurun_out/${TAG}_$c.err
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_${c}_launches.csv python bench.py --workload $c --no-cpu --steps 1 > gpurun_out/${TAG}_${c}_ncu_l.log 2>&1
done
EXTRA=smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,sm__cycles_active.avg,sm__cycles_elapsed.max
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:cloud_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG}_cloud -f python bench.py --workload config4 --no-cpu --steps 1 > gpurun_out/${TAG}_ncu_cloud.log 2>&1
# the spectral frame: primary, spawn, 8 waves, resolve; skip 3 warm-up frames (11 launches each), then capture waves 1 and 2
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:spectral_trace_kernel -s 27 -c 3 -o gpurun_out/prof_${TAG}_spectral -f python bench.py --workload config3 --no-cpu --steps 1 > gpurun_out/${TAG}_ncu_spectral.log 2>&1
cut -c1-400 gpurun_out/${TAG}_config3.json; echo; cut -c1-400 gpurun_out/${TAG}_config4.json; echo
