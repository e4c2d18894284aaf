#!/bin/bash
# second A/B call: MMA rates with 32-byte rows, ncu captures of the 256 -> 128 weight gradient (base / new) and of head.d (new)
O=gpurun_out
timeout 60 build/ab/mma_rate > $O/ab2_mma_rate.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 150 $NCU -k regex:wgrad_kernel -s 2 -c 1 -o $O/ab2_wg_b1c0_base -f python tools/ab_one.py build/ab/libbase.so wg_b1c0 3 > $O/ab2_ncu1.log 2>&1
timeout 150 $NCU -k regex:wgrad_kernel -s 2 -c 1 -o $O/ab2_wg_b1c0_new -f python tools/ab_one.py - wg_b1c0 3 > $O/ab2_ncu2.log 2>&1
timeout 150 $NCU -k regex:tapgemm_kernel -s 2 -c 1 -o $O/ab2_headd_new -f python tools/ab_one.py - headd 3 > $O/ab2_ncu3.log 2>&1
tail -12 $O/ab2_mma_rate.txt; tail -2 $O/ab2_ncu1.log $O/ab2_ncu2.log $O/ab2_ncu3.log; ls -la $O/*.ncu-rep
