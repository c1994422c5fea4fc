#!/bin/bash
# usage (inside gpurun): bash tests/ncu_final.sh <tag>   -- the three ncu passes whose summaries go under profiles/
tag=$1
# (a) every non-serial kernel with the full metric set (DRAM throughput against the 60 % gate, occupancy, stalls)
ncu --set full --clock-control none -k regex:'k_mf|rs_|k_analyze|k_filter|k_rc|k_adler|k_pack|k_unfilter' -c 80 -o gpurun_out/${tag}_mf -f python tests/prof_mf.py 256 > gpurun_out/${tag}_mf_ncu.log 2>&1
ncu -i gpurun_out/${tag}_mf.ncu-rep --page raw --csv > gpurun_out/${tag}_mf_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_mf.ncu-rep
# (b) DRAM bytes of one k_encode launch in the bench configuration
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_encode -c 1 --csv --log-file gpurun_out/${tag}_k_encode_traffic.csv python tests/prof_encode.py 1024 2 368 > gpurun_out/${tag}_traffic.log 2>&1
# (c) launch list of one bench step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --quick --no-cpu --steps 1 --warmup 1 > gpurun_out/${tag}_launches_bench.log 2>&1
tail -2 gpurun_out/${tag}_mf_ncu.log; tail -3 gpurun_out/${tag}_k_encode_traffic.csv | cut -c1-300
