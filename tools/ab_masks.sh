for cfg in "1 1" "1 0" "0 0" "0 1" "1 1" "0 0"; do
  set -- $cfg
  MULTIVAE_B200_MASK_D=$1 MULTIVAE_B200_MASK_H=$2 MV_BENCH_DUMP=gpurun_out/kt_$1$2.json timeout 300 python bench.py --no-cpu --no-check --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('MASK_D=$1 MASK_H=$2', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
python - <<'PY'
import json
for tag in ("11","10","00","01"):
    k=json.load(open(f"gpurun_out/kt_{tag}.json"))["kernels"]
    sel={n: round(v["ms"]/v["calls"]*1000) for n,v in k.items() if n in ("mv_tapgemm:head.d","mv_tapgemm:b3.c1","mv_tapgemm:b3.c0","mv_tapgemm:b3.c1d","mv_tapgemm:b3.c0d","mv_tapgemm:b2.c0","mv_tapgemm:b2.c1d","mv_tapgemm:b2.c1")}
    print(tag, sel)
PY
