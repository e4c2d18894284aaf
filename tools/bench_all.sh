#!/bin/bash
# All BASELINE configurations through bench.py on one GPU (results under gpurun_out/): tools/bench_all.sh <tag> [extra flags]
tag=$1; shift
for c in ns cfg2 cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 10 --warmup 3 "$@" > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err || echo "bench $c failed" >> gpurun_out/${tag}_bench_$c.err
done
