#!/bin/bash
# usage (inside gpurun): bash tests/ncu_capture.sh <tag> <kernel regex> <python driver + args...>
# source-counter capture of the first launch of one kernel; leaves <tag>_source.csv / <tag>_raw.csv under gpurun_out/
tag=$1; kern=$2; shift 2
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy \
    --clock-control none --import-source on -k regex:$kern -c 1 -o gpurun_out/$tag -f python "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
rm -f gpurun_out/$tag.ncu-rep
tail -2 gpurun_out/${tag}_ncu.log
