#!/bin/bash
# usage (inside gpurun): bash tests/ab_probe.sh "<probe specs>" lib1.so lib2.so ...   -- the same perf probe against several builds
specs=$1; shift
for so in "$@"; do
  echo "== $so"
  CSC_B200_LIB=$PWD/$so timeout 300 python tests/perf_probe.py $specs 2>&1 | cut -c1-360
done
