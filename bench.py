"""Benchmark of the hot path: one MMVAE+ PolyMNIST training step (5 modalities, K=10, DReG, ResNet
encoders/decoders, Adam) per "step" on synthetic batches, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = samples/s with the batch already resident in HBM; `e2e` =
the same metric through `BaseTrainer.step_batch` with pinned-host inputs copied H2D and the loss read
back D2H every step.  `roofline` = the dominant native kernel timed live with CUDA events on the
launching stream; `cpu_baseline` = the oracle port (CPU restatement of the reference) timed on this
box's host cores on a bounded sample of the same workload.  `--impl reference` times that CPU path
alone on the same config (the reference is pure Python/PyTorch: /root/reference does not travel, the
oracle port is its restatement — kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "MMVAE+ PolyMNIST K=10 train samples/sec"
M, K, L, LW, DIMS = 5, 10, 32, 32, (3, 28, 28)
D = 3 * 28 * 28
GFLOP_PER_SAMPLE = 166.15  # fwd+bwd, BASELINE.md section 3 (torch FlopCounterMode on the reference modules)


def workload_name(B):
    return (f"MMVAE+ PolyMNIST M=5 K=10 L=32+32 DReG laplace(0.75) beta=2.5 ResNet enc/dec, "
            f"per-GPU batch {B}, fwd+bwd+Adam")


def synthetic_batch(B, pinned=False):
    """U[0,1) images from per-modality generators."""
    data = {}
    for i in range(M):
        t = torch.rand(B, *DIMS, generator=torch.Generator().manual_seed(1000 + i))
        data[f"m{i}"] = t.pin_memory() if pinned else t
    return data


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md): nvidia-smi polled during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th is not None:
            self.th.join(timeout=10)
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU path: the oracle port (restatement of the reference's training step) — baseline only
# ------------------------------------------------------------------------------------------------
def cpu_port_runner(B, state_dict=None):
    """Returns (step_fn, nthreads): step_fn() runs forward + backward + Adam of the reference algorithm on
    CPU, fp32, with all host threads."""
    from oracle.port import elbo as E
    from oracle.port import nets as N
    nthreads = len(os.sched_getaffinity(0))
    torch.set_num_threads(nthreads)
    if state_dict is None:
        state_dict = north_star_model("cpu").state_dict()
    p = {k: v.detach().float().cpu().clone() for k, v in state_dict.items()}
    mods = [f"m{i}" for i in range(M)]
    for k, v in p.items():
        v.requires_grad_(not (k.startswith("mean_priors") or k == "logvars_priors.shared"))
    opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-3)
    data = synthetic_batch(B)
    enc = {m: (lambda x, m=m: N.encoder_resnet_mmnist(p, f"encoders.{m}.", x)) for m in mods}
    dec = {m: (lambda z, m=m: N.decoder_resnet_mmnist(p, f"decoders.{m}.", z)) for m in mods}
    g = torch.Generator().manual_seed(2000)

    def lap(shape):
        eps = torch.finfo(torch.float32).eps
        u = torch.empty(shape).uniform_(eps - 1, 1, generator=g)
        return -u.sign() * torch.log1p(-u.abs())

    def step():
        noise = {"u": {}, "w": {}, "prior": {}}
        for c in mods:
            noise["u"][c], noise["w"][c] = lap((K, B, L)), lap((K, B, LW))
            noise["prior"][c] = {r: lap((K, B, LW)) for r in mods if r != c}
        loss = E.mmvae_plus_forward(
            enc, dec, data, noise, K=K, latent_dim=L, style_dim=LW, beta=2.5, kind="laplace_with_softmax",
            loss="dreg_looser", dec_dist={m: "laplace" for m in mods}, dec_scale={m: 0.75 for m in mods},
            rescale={m: 1 for m in mods}, prior_mean={m: p[f"mean_priors.{m}"] for m in mods + ["shared"]},
            prior_logvar={m: p[f"logvars_priors.{m}"] for m in mods + ["shared"]})
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, nthreads


def time_cpu(B, steps, warmup):
    step, nthreads = cpu_port_runner(B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps, nthreads


# ------------------------------------------------------------------------------------------------
# GPU path
# ------------------------------------------------------------------------------------------------
def north_star_model(device):
    import multivae_b200 as mb
    from multivae_b200.nn import DecoderResnetMMNIST, EncoderResnetMMNIST
    mods = [f"m{i}" for i in range(M)]
    cfg = mb.MMVAEPlusConfig(
        n_modalities=M, input_dims={m: DIMS for m in mods}, K=K, latent_dim=L, modalities_specific_dim=LW, beta=2.5,
        prior_and_posterior_dist="laplace_with_softmax", decoders_dist={m: "laplace" for m in mods},
        decoder_dist_params={m: {"scale": 0.75} for m in mods}, learn_modality_prior=True, learn_shared_prior=False,
        loss="dreg_looser")
    torch.manual_seed(0)
    enc = {m: EncoderResnetMMNIST(LW, L) for m in mods}
    dec = {m: DecoderResnetMMNIST(L + LW) for m in mods}
    return mb.MMVAEPlus(cfg, enc, dec).to(device)


def elbo_rel_err(device, compute_dtype=torch.float32):
    """|loss_gpu - loss_reference| / |loss_reference| on the north-star golden produced by the REAL reference
    (tests/golden/elbo_ns_mmvaeplus_resnet.pt: ResNet encoders/decoders, K = 10, B = 4, recorded noise).  fp32: library
    networks + native ELBO kernels; bf16: the tcgen05 encoders/decoders (bf16 operands) + native ELBO kernels."""
    from tests.gpu_checks import rel, run_product
    out, _, rec = run_product("ns_mmvaeplus_resnet", device=device, compute_dtype=compute_dtype)
    return rel(out.loss.detach().cpu(), rec["loss"])


def run_gpu(args):
    import torch.distributed as dist

    import multivae_b200 as mb
    from multivae_b200 import _cabi
    from multivae_b200.trainer import BaseTrainer, BaseTrainerConfig

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: multivae_b200 has no CPU fallback "
                           "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    B = args.batch
    model = north_star_model(device)
    model.compute_dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    host = synthetic_batch(B, pinned=True)
    ds = mb.MultimodalBaseDataset(data=host)
    tcfg = BaseTrainerConfig(per_device_train_batch_size=B, learning_rate=1e-3, optimizer_cls="Adam",
                             world_size=world, rank=rank, local_rank=local_rank, use_cuda_graph=not args.no_graph)
    trainer = BaseTrainer(model, ds, training_config=tcfg)
    model.train()
    resident = mb.DatasetOutput(data={k: v.to(device) for k, v in host.items()})
    pinned = mb.DatasetOutput(data=host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host_ms[0] = (time.perf_counter() - t0) * 1e3 / steps   # host enqueue time per step (no sync inside)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident():
        trainer.step_batch(resident)

    last = {}

    def step_e2e():
        out = trainer.step_batch(pinned)  # H2D of the five pinned image tensors inside
        last["loss"] = float(out.loss_sum)  # D2H read of the step's result

    # untimed: W warm-up steps (eager), plus the CUDA-graph capture step and one replay when graphs are on
    for _ in range(args.warmup if args.no_graph else max(args.warmup, tcfg.graph_warmup_steps + 2)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    host_enqueue_ms = host_ms[0]
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel device times for the roofline (separate short pass so the events do not perturb `value`)
    trainer.step_batch(resident, allow_graph=False)   # un-timed eager pass: lazy module loading of every kernel variant
    timer = _cabi.KernelTimer()
    _cabi.set_timer(timer)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    nprof = max(1, min(3, args.steps))
    n0 = _cabi.launch_count()
    for _ in range(nprof):
        trainer.step_batch(resident, allow_graph=False)   # per-kernel events need host-launched kernels
    # kernels of this library per step (the same launches are what the CUDA graph replays in the timed region)
    launches = (_cabi.launch_count() - n0) // nprof * args.steps
    e1.record()
    summary = timer.summary()
    _cabi.set_timer(None)
    prof_ms = e0.elapsed_time(e1)
    # e2e leg
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    if rank != 0:
        return
    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    roof = roofline_from(summary, prof_ms, B, nprof, ms / args.steps)
    rel = rel16 = None
    if not args.no_check:
        try:
            rel = elbo_rel_err(device)
            rel16 = elbo_rel_err(device, torch.bfloat16)
        except Exception as e:  # the check must not hide the timing line
            rel = f"failed: {type(e).__name__}: {e}"
    cpu = None
    if not args.no_cpu:
        v, spstep, nthreads = time_cpu(args.cpu_batch, steps=args.cpu_steps, warmup=1)
        cpu = {"value": v, "unit": "samples/s", "cores": nthreads, "kind": "port",
               "sample": f"same workload at batch {args.cpu_batch}: 1 warm-up + {args.cpu_steps} timed steps of "
                         f"fwd+bwd+Adam, fp32, oracle/port (CPU restatement of the reference), {spstep:.2f} s/step",
               "cpu_count": os.cpu_count()}
    peaks = measured_peaks()
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_name(B), "per_gpu_batch": B, "global_batch": B * world,
                   "parallelism": f"dp{world}", "optimizer": "Adam lr=1e-3 (fp32 master weights)",
                   "l2": "no flush needed: each step streams > 10 GB of activations (>> 126 MB L2)",
                   "nn_backend": mb.nn.functional.backend_summary()},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": M * B * D * 4, "d2h_bytes_per_step": 4, "last_loss": last.get("loss")},
        "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
        "roofline": roof,
        "step_fraction_of_tensor_ceiling": value / world / (peaks["bf16_tflops_sustained"] * 1e3 / GFLOP_PER_SAMPLE),
        "elbo_rel_err": rel, "elbo_rel_err_bf16_tensor_path": rel16,
        "cpu_baseline": cpu,
    }
    EMIT(json.dumps(line))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        d["source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measured_traffic(name):
    """DRAM bytes per launch of a kernel from the committed `ncu --set full` capture (profiles/r1_ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f).get(name)
    return None if t is None else t["dram_read_bytes"] + t["dram_write_bytes"]


def roofline_from(summary, prof_ms, B, nprof, graph_ms_per_step):
    """Roofline of the dominant native kernel from the live per-entry-point CUDA-event times.  Shares are kernel time per
    step (host-launched pass with events) over the CUDA-graph step time of the timed region."""
    from multivae_b200 import roofline as R
    peaks = measured_peaks()
    if not summary:
        return None
    top = max(summary.items(), key=lambda kv: kv[1][1])
    name, (calls, total_ms) = top
    info = R.describe(name, B=B, M=M, K=K, D=D, L=L, LW=LW)
    per_launch_s = total_ms / calls / 1e3
    if info["bound"] == "hbm":
        achieved = info["work"] / per_launch_s / 1e9
        peak, unit = peaks["hbm_gbs"], "GB/s"
    else:
        achieved = info["work"] / per_launch_s / 1e12
        peak, unit = peaks["bf16_tflops_sustained"], "TFLOP/s"
    shares = {k: round(v[1] / nprof / graph_ms_per_step, 4) for k, v in sorted(summary.items(), key=lambda kv: -kv[1][1])[:8]}
    # the HBM-bound fused ELBO kernels next to it (north_star asks for both rooflines)
    elbo = {}
    for sym in ("mv_moe_lpx_fwd_multi", "mv_moe_lpx_bwd_multi", "mv_moe_lpx_fwd", "mv_moe_lpx_bwd", "mv_moe_lw_fwd"):
        if sym in summary:
            c, t = summary[sym]
            w = R.describe(sym, B=B, M=M, K=K, D=D, L=L, LW=LW)["work"]
            elbo[sym] = {"avg_launch_us": t / c * 1e3, "achieved_GBps": w / (t / c / 1e3) / 1e9,
                         "frac_of_measured_hbm": w / (t / c / 1e3) / 1e9 / peaks["hbm_gbs"], "bytes_per_launch": w}
    # all tensor-core launches together: useful flops / summed kernel time
    tflops = tms = 0.0
    for k, (c, t) in summary.items():
        inf = R.describe(k, B=B, M=M, K=K, D=D, L=L, LW=LW)
        if inf["bound"] == "tensor":
            tflops += inf["work"] * c
            tms += t
    tensor_all = {"useful_TFLOPs_per_s": tflops / (tms / 1e3) / 1e12 if tms else None,
                  "frac_of_sustained_peak": tflops / (tms / 1e3) / 1e12 / peaks["bf16_tflops_sustained"] if tms else None,
                  "share_of_step": tms / nprof / graph_ms_per_step}
    if os.environ.get("MV_BENCH_DUMP"):
        with open(os.environ["MV_BENCH_DUMP"], "w") as f:
            json.dump({"prof_ms": prof_ms, "kernels": {k: {"calls": v[0], "ms": v[1]} for k, v in summary.items()}}, f, indent=1)
    return {"kernel": name, "bound": info["bound"], "achieved": achieved, "peak": peak, "unit": unit,
            "frac": achieved / peak, "peak_source": peaks["source"], "traffic": measured_traffic(name),
            "algorithmic_work_per_launch": info["work"], "avg_launch_us": per_launch_s * 1e6, "launches": calls,
            "share_of_step": total_ms / nprof / graph_ms_per_step, "shares": shares, "elbo_kernels": elbo, "tensor_kernels_total": tensor_all}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, spstep, nthreads = time_cpu(args.cpu_batch, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch),
                   "sample": f"each step = the same training step at batch {args.cpu_batch} on the host CPU"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": nthreads, "kind": "port",
                         "sample": f"batch {args.cpu_batch} per step, {args.steps} timed steps, fp32, all host threads"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    EMIT(json.dumps(line))


def _claim_stdout():
    """Everything but the final JSON line goes to stderr (NCCL and torch print banners on fd 1): returns a writer for the
    real stdout."""
    real = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    sys.stdout = sys.stderr

    def emit(line):
        os.write(real, (line + "\n").encode())
    return emit


EMIT = print


def main():
    global EMIT
    EMIT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun when called bare with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu(args)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
