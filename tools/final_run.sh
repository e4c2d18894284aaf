#!/bin/bash
# End-of-round evidence run on ONE GPU (under gpurun): tests, smoke, every BASELINE configuration through bench.py (with the CPU
# baseline and the library-eager arm), the reference arm, ncu launch list of the north-star step.  Outputs under gpurun_out/.
O=gpurun_out; T=${1:-final}
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
MV_BENCH_DUMP=$O/${T}_kernel_times_ns.json timeout 900 python bench.py --steps 20 --warmup 5 --torch-eager > $O/${T}_bench_ns.json 2> $O/${T}_bench_ns.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/${T}_bench_ns_reference.json 2>> $O/${T}_bench_ns.err
for c in cfg2 cfg3 cfg4 cfg5; do
  MV_BENCH_DUMP=$O/${T}_kernel_times_$c.json timeout 600 python bench.py --config $c --steps 20 --warmup 5 --torch-eager > $O/${T}_bench_$c.json 2> $O/${T}_bench_$c.err
done
ls -la $O | tail -20
