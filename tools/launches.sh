#!/bin/bash
# usage: tools/launches.sh <n_agents> <steps> <tag> [extra probe args]  -> gpurun_out/launches_<tag>.csv (run on the GPU box)
N=$1; S=$2; TAG=$3; shift 3
python tools/probe.py --n $N --iso hard --steps $S --reps 1 "$@" > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:seirm -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python tools/probe.py --n $N --iso hard --steps $S --reps 1 "$@" > gpurun_out/ncu_$TAG.log 2>&1
