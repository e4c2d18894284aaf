"""Parity report on the GPU: every golden case through the product path, errors printed (no asserts).
    python tools/parity_report.py [case ...]  ->  one JSON line per case (also appended to gpurun_out/parity_report.jsonl)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle.cases import CASES  # noqa: E402
from oracle.replay import run_port  # noqa: E402
from tests.gpu_checks import rel, run_product  # noqa: E402


def report(name, dtype=None):
    out, model, rec = run_product(name, compute_dtype=dtype)
    r = {"case": name, "dtype": "bf16" if dtype is not None else "f32", "loss": float(out.loss), "ref": float(rec["loss"]),
         "loss_rel": rel(out.loss.detach().cpu(), rec["loss"])}
    if "lws" in rec:
        lw = model._last["lw"].cpu()
        e = 0.0
        for i, m in enumerate(rec["lws"]):
            d = (lw[i] - rec["lws"][m]).abs()
            e = max(e, float((d / rec["lws"][m].abs().clamp(min=1.0)).max()))
        r["lw_max_rel"] = e
    for k, v in rec["metrics"].items():
        got = out.metrics[k]
        got = got.detach().cpu() if torch.is_tensor(got) else got
        r.setdefault("metrics_rel", {})[k] = rel(got, v)
    _, _, _, pp = run_port(CASES[name], rec)
    worst, wk, num, den = 0.0, None, 0.0, 0.0
    for k, p in model.named_parameters():
        if rec["grads"][k] is None or p.grad is None:
            continue
        ref = pp[k].grad
        pg = p.grad.detach().float().cpu()
        e = float((pg - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
        if e > worst:
            worst, wk = e, k
        num += float((pg - ref).double().pow(2).sum())
        den += float(ref.double().pow(2).sum())
    r.update(grad_max_rel=worst, grad_worst=wk, grad_rel_l2=(num / max(den, 1e-30)) ** 0.5)
    return r


if __name__ == "__main__":
    names = sys.argv[1:] or sorted(CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
        for n in names:
            for dt in ([None, torch.bfloat16] if "arch" in CASES[n] or n.startswith("cfg") else [None]):
                try:
                    r = report(n, dt)
                except Exception as e:  # keep going: this is a report
                    r = {"case": n, "dtype": str(dt), "error": f"{type(e).__name__}: {e}"}
                line = json.dumps(r)
                print(line, flush=True)
                f.write(line + "\n")
