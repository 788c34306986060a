#!/bin/bash
# usage: ncu_one.sh <name> <kernel regex> <bench_config args...>
name=$1; shift; kre=$1; shift
ncu --set full --clock-control none --import-source on -k regex:$kre -c 1 -o gpurun_out/$name -f python tools/bench_config.py "$@" > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log
