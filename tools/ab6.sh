#!/bin/bash
O=gpurun_out
for L in multivae_b200/libmultivae_b200.so build/ab/libexp_direct.so; do
  echo "=== $L"
  timeout 100 python tools/ab_bench.py $L 10 2>&1 | grep -E "head.d|e.img"
  timeout 100 python tools/conv3_bench.py 10 $L 2>&1 | grep -E "c1d"
done > $O/ab6.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 2 -c 1 -o $O/ab6_headd_lean -f python tools/ab_one.py - headd 3 > $O/ab6_ncu.log 2>&1
cat $O/ab6.log
