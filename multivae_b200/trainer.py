"""Host side of the training step: a mirror of the reference's `BaseTrainer.train_step` /
`_optimizers_step` / DDP wrap (/root/reference/src/multivae/trainers/base/base_trainer.py:350-361,
682-750, 92-117, 171-219) for the five hot-path models, built B200-first:

  * one process per GPU (`LOCAL_RANK`/`RANK`/`WORLD_SIZE` from the environment, like
    base_trainer_config.py:80-100); the batch is sharded by rank exactly like
    `DistributedSampler(num_replicas, rank)` (its default shuffle=True: a new permutation every epoch); parameters are replicated;
  * gradients live in ONE flat fp32 buffer per model (every `p.grad` is a view into it), so the
    data-parallel exchange is one NCCL all-reduce (mean) over NVLink per step instead of DDP's 25 MB
    buckets (payloads are 6-89 MB) and `zero_grad` is one memset;
  * no per-step `torch.cuda.empty_cache()` and no per-step `.item()`: `loss_sum` is accumulated on
    the device and read back once per epoch (the NaN guard keeps the reference's `ArithmeticError`).

Out of scope here (the reference's Python stays as is): callbacks, checkpoints, schedulers, image logging.
"""
import os
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.distributed as dist

from .containers import DatasetOutput


@dataclass
class BaseTrainerConfig:
    """The subset of base_trainer_config.py:10-152 the training step reads."""
    per_device_train_batch_size: int = 64
    per_device_eval_batch_size: int = 64
    num_epochs: int = 100
    optimizer_cls: str = "Adam"
    optimizer_params: Optional[dict] = None
    learning_rate: float = 1e-4
    no_cuda: bool = False
    world_size: int = -1
    local_rank: int = -1
    rank: int = -1
    dist_backend: str = "nccl"
    master_addr: str = "127.0.0.1"
    master_port: str = "12345"
    gradient_clipping_max_norm: Optional[float] = None
    # B200 addition: replay the whole step (forward, backward, all-reduce, optimizer) as ONE CUDA graph once the
    # batch shape has been seen `graph_warmup_steps` times — the step is ~10^3 kernel launches and host-bound otherwise
    use_cuda_graph: bool = False
    graph_warmup_steps: int = 3
    # batch order: the reference builds DataLoader(shuffle=True) / DistributedSampler(shuffle=True) (base_trainer.py:199-210)
    shuffle: bool = True
    seed: int = 0
    # reduce the decoders' gradient slice on a side stream while the encoders' backward still runs (multi-GPU only)
    overlap_allreduce: bool = True
    beta_schedule: Optional[list] = field(default=None)

    def __post_init__(self):
        env = os.environ
        if self.world_size == -1 and "WORLD_SIZE" in env:
            self.world_size = int(env["WORLD_SIZE"])
        if self.rank == -1 and "RANK" in env:
            self.rank = int(env["RANK"])
        if self.local_rank == -1 and "LOCAL_RANK" in env:
            self.local_rank = int(env["LOCAL_RANK"])
        if not hasattr(torch.optim, self.optimizer_cls):
            raise AttributeError(f"Unable to import `{self.optimizer_cls}` optimizer from 'torch.optim'.")


def shard_indices(n, world_size, rank, shuffle=False, seed=0, epoch=0):
    """Indices of `DistributedSampler(dataset, num_replicas=world_size, rank=rank, shuffle=shuffle, seed=seed,
    drop_last=False)` after `set_epoch(epoch)`: a `randperm` from a generator seeded with seed + epoch (shuffle), padded to a
    multiple of world_size by wrapping, then rank::world_size — torch/utils/data/distributed.py, what the reference's
    trainer builds with the sampler's default shuffle=True (base_trainer.py:198-201)."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g).tolist()
    else:
        idx = list(range(n))
    total = (n + world_size - 1) // world_size * world_size
    pad = total - n
    if pad <= len(idx):
        idx += idx[:pad]
    else:
        idx += (idx * ((pad + len(idx) - 1) // len(idx)))[:pad]
    return idx[rank:total:world_size]


class StagedBatch(DatasetOutput):
    """A batch whose host->device copy was started ahead of time by `BaseTrainer.prefetch` (device tensors in `data`, the event
    `ready` fires when the copy has landed, `slot` is the staging set to hand back)."""


class FlatGrads:
    """All parameter gradients as views of one flat fp32 buffer (the all-reduce bucket)."""

    def __init__(self, params, first=()):
        """`first`: parameters laid out at the front of the buffer (the decoders: their gradients are final before the encoders'
        backward starts, so that slice can be reduced early); `split` = number of elements of that prefix."""
        first_ids = {id(p) for p in first}
        ps = [p for p in params if p.requires_grad]
        self.params = [p for p in ps if id(p) in first_ids] + [p for p in ps if id(p) not in first_ids]
        self.split = sum(p.numel() for p in self.params if id(p) in first_ids)
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            assert p.dtype == torch.float32, "master parameters are fp32 (reference is fp32 end to end)"
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def rebind(self):
        """Autograd may replace .grad when it was set to None by user code; restore the views."""
        o = 0
        for p in self.params:
            v = self.flat[o:o + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
            o += p.numel()

    def allreduce_mean(self, group=None, lo=0, hi=None):
        """DDP semantics: grad = (1/W) * sum_r grad_r  (base_trainer.py:116-117), on elements [lo, hi) of the buffer."""
        w = dist.get_world_size(group)
        if w == 1:
            return
        t = self.flat if (lo == 0 and hi is None) else self.flat[lo:hi]
        if t.numel() == 0:
            return
        if self.flat.is_cuda:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
        else:  # gloo (CPU tests) has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t.div_(w)


class BaseTrainer:
    """`BaseTrainer(model, train_dataset, training_config=...)`; `.train_step(epoch)` returns
    `(epoch_loss, epoch_metrics)` like the reference (epoch_loss = sum of loss_sum / len(train_dataset),
    metrics averaged over batches)."""

    def __init__(self, model, train_dataset, eval_dataset=None, training_config=None, callbacks=None,
                 checkpoint=None):
        self.training_config = training_config or BaseTrainerConfig()
        cfg = self.training_config
        self.model = model
        self.train_dataset = train_dataset
        self.eval_dataset = eval_dataset
        self.world_size = max(cfg.world_size, 1)
        self.rank = max(cfg.rank, 0)
        self.local_rank = max(cfg.local_rank, 0)
        self.distributed = cfg.world_size > 1
        self.device = self._setup_devices()
        self.model.to(self.device)
        self.model.device = self.device
        decs = getattr(self.model, "decoders", None)
        self.flat = FlatGrads(self.model.parameters(), first=list(decs.parameters()) if decs is not None else ())
        self._comm_stream = None
        self._graphs = {}
        self.set_optimizer()
        self._broadcast_parameters()

    # base_trainer.py:171-193
    def _setup_devices(self):
        cfg = self.training_config
        if cfg.no_cuda or not torch.cuda.is_available():
            device = torch.device("cpu")
        else:
            torch.cuda.set_device(self.local_rank)
            device = torch.device("cuda", self.local_rank)
        if self.distributed and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", cfg.master_addr)
            os.environ.setdefault("MASTER_PORT", cfg.master_port)
            backend = cfg.dist_backend if device.type == "cuda" else "gloo"
            dist.init_process_group(backend=backend, init_method="env://", world_size=self.world_size,
                                    rank=self.rank)
        return device

    def _broadcast_parameters(self):
        """DDP's initial rank-0 parameter/buffer broadcast."""
        if self.distributed:
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=0)

    # base_trainer.py:230-246
    def set_optimizer(self):
        cfg = self.training_config
        cls = getattr(torch.optim, cfg.optimizer_cls)
        kw = dict(cfg.optimizer_params or {})
        if self.device.type == "cuda" and cfg.optimizer_cls in ("Adam", "AdamW", "SGD"):
            kw.setdefault("fused", True)
        if self.device.type == "cuda" and cfg.use_cuda_graph and cfg.optimizer_cls in ("Adam", "AdamW"):
            kw.setdefault("capturable", True)  # step counter on the device
        self.optimizer = cls([p for p in self.model.parameters() if p.requires_grad], lr=cfg.learning_rate, **kw)

    def _batches_of(self, dataset, idx, bs):
        for i in range(0, len(idx), bs):
            sel = idx[i:i + bs]
            contiguous = sel == list(range(sel[0], sel[0] + len(sel)))
            take = (lambda t: t[sel[0]:sel[0] + len(sel)]) if contiguous else (lambda t: t[torch.as_tensor(sel)])
            out = DatasetOutput(data={m: take(t) for m, t in dataset.data.items()})
            if hasattr(dataset, "masks"):
                out["masks"] = {m: take(t) for m, t in dataset.masks.items()}
            yield out

    def local_eval_batches(self):
        """This rank's batches of the evaluation set (base_trainer.py:212-219: no shuffling in a single process; the
        DistributedSampler's default permutation of seed + 0 when distributed)."""
        n = len(self.eval_dataset)
        idx = (shard_indices(n, self.world_size, self.rank, shuffle=True, seed=0, epoch=0) if self.distributed else list(range(n)))
        return self._batches_of(self.eval_dataset, idx, self.training_config.per_device_eval_batch_size)

    def local_batches(self, epoch=0):
        """This rank's batches of the dataset for `epoch`: a fresh permutation per epoch (single process: the DataLoader's
        shuffle=True; distributed: the DistributedSampler permutation of seed + epoch, sharded rank::world)."""
        n = len(self.train_dataset)
        cfg = self.training_config
        idx = shard_indices(n, self.world_size if self.distributed else 1, self.rank if self.distributed else 0,
                            shuffle=cfg.shuffle, seed=cfg.seed, epoch=epoch)
        return self._batches_of(self.train_dataset, idx, cfg.per_device_train_batch_size)

    def prefetch(self, inputs):
        """Start the host->device copy of a (pinned) batch on a copy stream into one of two persistent staging sets and return a
        StagedBatch for `step_batch`: called for batch i+1 right after batch i was launched, the PCIe transfer overlaps batch i's
        compute instead of sitting between two steps (double buffering; the reference's DataLoader does the same with
        `pin_memory` + non_blocking copies).  Batches with masks, CPU runs and non-tensor payloads are returned unchanged."""
        if self.device.type != "cuda" or hasattr(inputs, "masks") or isinstance(inputs, StagedBatch):
            return inputs
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stages, self._stage_free, self._stage_next = {}, {}, 0
        sig = tuple((m, tuple(t.shape), t.dtype) for m, t in inputs.data.items())
        slot = self._stage_next
        self._stage_next ^= 1
        key = (sig, slot)
        if key not in self._stages:
            self._stages[key] = {m: torch.empty(t.shape, dtype=t.dtype, device=self.device) for m, t in inputs.data.items()}
        cs = self._copy_stream
        if key in self._stage_free:
            cs.wait_event(self._stage_free[key])   # the step that consumed this staging set last has copied it out
        with torch.cuda.stream(cs):
            for m, t in inputs.data.items():
                self._stages[key][m].copy_(t, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cs)
        out = StagedBatch(data=self._stages[key])
        out["ready"], out["slot"] = ready, key
        return out

    def _consume_staged(self, inputs):
        """Make the current stream wait for a staged batch's copy; returns the device-resident batch."""
        torch.cuda.current_stream().wait_event(inputs.ready)
        return inputs

    def _release_staged(self, inputs):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._stage_free[inputs.slot] = ev

    def _to_device(self, inputs):
        mv = lambda t: t.to(self.device, non_blocking=True)  # noqa: E731
        out = DatasetOutput(data={m: mv(t) for m, t in inputs.data.items()})
        if hasattr(inputs, "masks"):
            out["masks"] = {m: mv(t) for m, t in inputs.masks.items()}
        return out

    # base_trainer.py:350-361
    def _optimizers_step(self, model_output):
        from .nn import resnet_native as RN
        self.flat.zero()
        early = self._arm_early_allreduce()
        with RN.direct_grads(True):   # the native stacks may add their weight gradients straight into the flat buffer's views
            model_output.loss.backward()
        self.flat.rebind()
        if self.distributed:
            if early is not None and early["fired"]:
                # the decoder slice was reduced on the side stream while the encoders' backward ran: join, then the rest
                early["handle"].remove()
                torch.cuda.current_stream().wait_stream(self._comm_stream)
                self.flat.allreduce_mean(lo=self.flat.split)
            else:
                if early is not None:
                    early["handle"].remove()
                self.flat.allreduce_mean()
        # parameters that received NO gradient (modalities missing from the whole batch) must not move: the reference's
        # zero_grad() leaves their .grad at None and the optimizer skips them (no stale-momentum / weight-decay update)
        skipped = []
        for m in getattr(self.model, "_unused_modalities", None) or ():
            for mod in (self.model.encoders[m], self.model.decoders[m]):
                for p in mod.parameters():
                    if p.grad is not None:
                        skipped.append(p)
                        p.grad = None
        self.optimizer.step()
        if skipped:
            self.flat.rebind()

    def _arm_early_allreduce(self):
        """Overlap of the gradient exchange with the backward pass: the decoders' gradients (the front slice of the flat buffer,
        the bulk of the payload) are final as soon as every decoder's backward has produced the gradient of its INPUT, long
        before the encoders' backward ends.  A multi-tensor gradient hook on the decoder inputs (the model lists them in
        `_decoder_inputs`; the native stacks write their weight gradients into the flat buffer inside their backward) launches the
        all-reduce of that slice on a side stream; the step joins it before reducing the rest.  Works inside CUDA-graph capture
        (the side stream forks from and rejoins the capturing stream).  Returns None when not applicable."""
        zs = [z for z in (getattr(self.model, "_decoder_inputs", None) or []) if z.requires_grad]
        if not (self.distributed and self.flat.flat.is_cuda and self.flat.split > 0 and zs and self.training_config.overlap_allreduce):
            return None
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        state = {"fired": False}

        def fire(_grads):
            if state["fired"]:
                return
            state["fired"] = True
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            for st in (getattr(self.model, "_dec_streams", None) or {}).values():   # decoders that ran on their own streams
                self._comm_stream.wait_stream(st)
            with torch.cuda.stream(self._comm_stream):
                self.flat.allreduce_mean(hi=self.flat.split)

        state["handle"] = torch.autograd.graph.register_multi_grad_hook(tuple(zs), fire, mode="all")
        return state

    def _eager_step(self, inputs, epoch, batch_ratio, **extra):
        sched = self.training_config.beta_schedule
        beta_epoch = sched[epoch - 1] if sched is not None else 1
        out = self.model(inputs, epoch=epoch, dataset_size=len(self.train_dataset), uses_ddp=self.distributed,
                         batch_ratio=batch_ratio, beta=beta_epoch, **extra)
        self._optimizers_step(out)
        return out

    def step_batch(self, inputs, epoch=1, batch_ratio=0.0, n_batches=1, allow_graph=True):
        """One batch of train_step (base_trainer.py:704-728) without the host read-back."""
        cfg = self.training_config
        staged = isinstance(inputs, StagedBatch)
        if staged:
            self._consume_staged(inputs)
        if not (cfg.use_cuda_graph and allow_graph and self.device.type == "cuda"):
            out = self._eager_step(self._to_device(inputs), epoch, batch_ratio)
        else:
            out = self._graph_step(inputs, epoch, batch_ratio)
        if staged:
            # eager: the step's kernels read the staging tensors directly; graph: they were copied into the static inputs
            self._release_staged(inputs)
        return out

    # ---- CUDA-graph replay of the step ------------------------------------------------------------
    def _graph_step(self, inputs, epoch, batch_ratio):
        """Eager for the first `graph_warmup_steps` occurrences of a batch signature, then capture once and replay.
        Per-step Python scalars the models read (MVAE's KL warm-up weight) live in device scalars refreshed by
        `model.prepare_step` outside the graph, so one capture serves every epoch; models whose step draws host-side
        randomness (`graph_safe` False: MVAE with k random subsets) and incomplete batches (masks: data-dependent host
        decisions) always run eagerly."""
        has_masks = hasattr(inputs, "masks")
        if has_masks or not getattr(self.model, "graph_safe", True):
            return self._eager_step(self._to_device(inputs), epoch, batch_ratio)
        sig = tuple((m, tuple(t.shape), t.dtype) for m, t in inputs.data.items())
        st = self._graphs.setdefault(sig, {"seen": 0})
        extra = {"beta_on_device": True} if hasattr(self.model, "_beta_dev") else {}
        self.model.prepare_step(epoch, batch_ratio)
        # Warm-up and capture both run on ONE side stream: autograd binds each parameter's AccumulateGrad node to the stream of
        # the first backward that uses it, and a node bound to the default stream makes the captured backward synchronise with
        # the (non-capturing) default stream, which invalidates the capture.
        if getattr(self, "_graph_stream", None) is None:
            self._graph_stream = torch.cuda.Stream(device=self.device)
        side = self._graph_stream
        if "graph" not in st:
            if st["seen"] < self.training_config.graph_warmup_steps:
                st["seen"] += 1
                dev_inputs = self._to_device(inputs)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    out = self._eager_step(dev_inputs, epoch, batch_ratio, **extra)
                torch.cuda.current_stream().wait_stream(side)
                return out
            static = DatasetOutput(data={m: torch.empty(t.shape, dtype=t.dtype, device=self.device)
                                         for m, t in inputs.data.items()})
            for m, t in inputs.data.items():
                static.data[m].copy_(t, non_blocking=True)
            torch.cuda.synchronize()
            torch.cuda.empty_cache()   # hand the eager pool's blocks to the graph's private pool
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                out = self._eager_step(static, epoch, batch_ratio, **extra)
            st.update(graph=graph, static=static, out=out)
        else:
            for m, t in inputs.data.items():
                st["static"].data[m].copy_(t, non_blocking=True)
        st["graph"].replay()
        return st["out"]

    # base_trainer.py:682-750
    def train_step(self, epoch):
        self.model.train()
        batches = list(self.local_batches(epoch))
        epoch_loss = torch.zeros((), device=self.device, dtype=torch.float64)
        metrics = {}
        nxt = self.prefetch(batches[0]) if batches else None
        for i, inputs in enumerate(batches):
            cur = nxt
            out = self.step_batch(cur, epoch=epoch, batch_ratio=i / len(batches), n_batches=len(batches))
            if i + 1 < len(batches):
                nxt = self.prefetch(batches[i + 1])   # the next batch's H2D copy overlaps this step's compute
            loss = out.loss_sum if hasattr(out, "loss_sum") else out.loss
            epoch_loss += loss.detach().double()
            for k, v in out.metrics.items():
                v = v.detach().double() if torch.is_tensor(v) else torch.tensor(float(v), device=self.device, dtype=torch.float64)
                metrics[k] = metrics.get(k, 0) + v
        total = float(epoch_loss)  # the one device->host sync of the epoch
        if total != total:
            raise ArithmeticError("NaN detected in train loss")
        self.model.update()
        metrics = {k: float(v) / len(batches) for k, v in metrics.items()}
        return total / len(self.train_dataset), metrics

    # base_trainer.py:618-680
    def eval_step(self, epoch):
        """Evaluation pass: `model(inputs, ..., use_mean_embedding=True)` under no_grad over the evaluation set; returns
        (sum of loss_sum / len(eval_dataset), metrics averaged over batches) like the reference.  The loss is accumulated on the
        device and read back once (the NaN guard keeps the reference's ArithmeticError)."""
        if self.eval_dataset is None:
            raise AttributeError("eval_step needs an eval_dataset")
        self.model.eval()
        batches = list(self.local_eval_batches())
        epoch_loss = torch.zeros((), device=self.device, dtype=torch.float64)
        metrics = {}
        for inputs in batches:
            with torch.no_grad():
                out = self.model(self._to_device(inputs), epoch=epoch, dataset_size=len(self.eval_dataset),
                                 uses_ddp=self.distributed, use_mean_embedding=True)
            loss = out.loss_sum if hasattr(out, "loss_sum") else out.loss
            epoch_loss += loss.detach().double()
            for k, v in out.metrics.items():
                v = v.detach().double() if torch.is_tensor(v) else torch.tensor(float(v), device=self.device, dtype=torch.float64)
                metrics[k] = metrics.get(k, 0) + v
        total = float(epoch_loss)
        if total != total:
            raise ArithmeticError("NaN detected in eval loss")
        return total / len(self.eval_dataset), {k: float(v) / len(batches) for k, v in metrics.items()}
