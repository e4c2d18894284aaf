#!/bin/bash
# ncu evidence of round 2 (run under gpurun, one GPU): launch lists of whole steps + full captures of the kernels DESIGN.md cites
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# launch lists (per-launch times are cold-cache and serialised: the SHARES are what counts); step 1 = lazy loading, step 2 is listed
for c in ns cfg3 cfg5; do
  timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/r2_launches_$c.csv python tools/ncu_step.py $c 2 > $O/r2_launches_$c.log 2>&1
done
# full captures (3 launches each, skipping the first step's launches of that kernel)
timeout 600 $NCU --set full --import-source on -k regex:gemm_kernel -s 40 -c 6 -o $O/r2_ncu_gemm_cfg5 -f python tools/ncu_step.py cfg5 3 > $O/r2_ncu_gemm_cfg5.log 2>&1
timeout 600 $NCU --set full --import-source on -k "regex:tapgemm_kernel<16" -s 5 -c 2 -o $O/r2_ncu_headd -f python tools/ncu_step.py ns 2 > $O/r2_ncu_headd.log 2>&1
timeout 600 $NCU --set full --import-source on -k "regex:conv3_kernel<19" -s 5 -c 2 -o $O/r2_ncu_b3c1 -f python tools/ncu_step.py ns 2 > $O/r2_ncu_b3c1.log 2>&1
timeout 600 $NCU --set full --import-source on -k "regex:col2im|im2col" -s 20 -c 6 -o $O/r2_ncu_gather_cfg3 -f python tools/ncu_step.py cfg3 3 > $O/r2_ncu_gather_cfg3.log 2>&1
ls -la $O/*.ncu-rep
