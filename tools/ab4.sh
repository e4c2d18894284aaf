#!/bin/bash
# ncu source-level captures of the dominant three-tap convolutions + a north-star bench line after the head.d / wgrad changes
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 150 $NCU -k regex:conv3_kernel -s 2 -c 1 -o $O/ab4_c3_c1 -f python tools/ab_one.py - c3_c1 3 > $O/ab4_ncu1.log 2>&1
timeout 150 $NCU -k regex:conv3_kernel -s 2 -c 1 -o $O/ab4_c3_c0 -f python tools/ab_one.py - c3_c0 3 > $O/ab4_ncu2.log 2>&1
timeout 150 $NCU -k regex:conv3_kernel -s 2 -c 1 -o $O/ab4_c3_c1d -f python tools/ab_one.py - c3_c1d 3 > $O/ab4_ncu3.log 2>&1
MV_BENCH_DUMP=$O/ab4_kernel_times_ns.json timeout 600 python bench.py --steps 20 --warmup 5 > $O/ab4_bench_ns.json 2> $O/ab4_bench_ns.err
tail -c 600 $O/ab4_bench_ns.json; tail -3 $O/ab4_bench_ns.err; ls -la $O/ab4*
