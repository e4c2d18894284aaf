timeout 600 python -m pytest tests/test_gpu_resnet_decoder.py tests/test_gpu_resnet_encoder.py -x -q 2>&1 | tail -3
for v in 1 0 1 0; do
  MULTIVAE_B200_SPLIT_C0D=$v MV_BENCH_DUMP=gpurun_out/kt_sp$v.json timeout 300 python bench.py --no-cpu --no-check --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('SPLIT_C0D=$v', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
python - <<'PY'
import json
for tag in ("1","0"):
    k=json.load(open(f"gpurun_out/kt_sp{tag}.json"))["kernels"]
    sel={n.split(':')[1]: (v["calls"]//3, round(v["ms"]/3,2)) for n,v in k.items() if "c0d" in n}
    print("SPLIT="+tag, sel)
PY
