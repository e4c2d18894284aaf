"""Key fields of a bench.py JSON line: python tools/print_bench.py gpurun_out/bench.json"""
import json, sys
txt = open(sys.argv[1]).read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("value", round(d["value"], 1), d["unit"], "| ms/step", round(d["ms_per_step"], 2), "| e2e", round(d["e2e"]["value"], 1), "| n_gpus", d["n_gpus"],
      "| launches", d.get("gpu_launches"), "| clocks", d.get("clocks"))
print("elbo_rel_err", d.get("elbo_rel_err"), d.get("elbo_rel_err_bf16_tensor_path"), "| cpu_baseline", (d.get("cpu_baseline") or {}).get("value"))
r = d.get("roofline") or {}
print("roofline", r.get("kernel"), r.get("bound"), "achieved", round(r.get("achieved", 0), 1), r.get("unit"), "frac", round(r.get("frac", 0), 3), "traffic", r.get("traffic"))
print("elbo kernels", {k: (round(v["avg_launch_us"]), round(v["frac_of_measured_hbm"], 3)) for k, v in (r.get("elbo_kernels") or {}).items()})
print("tensor total", r.get("tensor_kernels_total"))
