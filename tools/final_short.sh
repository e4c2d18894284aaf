#!/bin/bash
# Short end-of-round check on ONE GPU: all GPU tests, smoke, the north-star bench line and one small configuration.
O=gpurun_out; T=${1:-fin}
timeout 600 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
MV_BENCH_DUMP=$O/${T}_kernel_times_ns.json timeout 600 python bench.py --steps 20 --warmup 5 > $O/${T}_bench_ns.json 2> $O/${T}_bench_ns.err
MV_BENCH_DUMP=$O/${T}_kernel_times_cfg3.json timeout 300 python bench.py --config cfg3 --steps 20 --warmup 5 > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
python - <<PY
import json
for c in ("ns", "cfg3"):
    d = json.load(open("$O/${T}_bench_%s.json" % c))
    print(c, round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d["clocks"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
PY
