timeout 600 python -m pytest tests/test_gpu_tapgemm.py tests/test_gpu_resnet_decoder.py tests/test_gpu_resnet_encoder.py -x -q 2>&1 | tail -3
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export MV_TG_NO_SEQ=1; else unset MV_TG_NO_SEQ; fi
  MV_BENCH_DUMP=gpurun_out/kt_seq$v.json timeout 300 python bench.py --no-cpu --no-check --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NO_SEQ=$v', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
python - <<'PY'
import json
for tag in ("0","1"):
    k=json.load(open(f"gpurun_out/kt_seq{tag}.json"))["kernels"]
    sel={n.split(':')[1]: round(v["ms"]/v["calls"]*1000) for n,v in k.items() if n.startswith("mv_tapgemm:b1") or n.startswith("mv_tapgemm:b2.c0d") or n.startswith("mv_tapgemm:b2.sc") or n.startswith("mv_tapgemm:e3") or n.startswith("mv_tapgemm:e2.c1")}
    print("NO_SEQ="+tag, sel)
PY
