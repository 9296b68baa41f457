"""Training-step and sampling engines: the loop bodies of the reference's train.py / inference.py
on CUDA streams and graphs.

TrainEngine.step is what `train_one_epoch` does per batch (train.py:829-867) -- fp16 forward/backward with
a dynamically scaled loss (autocast + GradScaler, train.py:853-867), unscale + inf check, global-norm clip,
AdamW, scaler update -- with two differences that are deliberate and documented in DESIGN.md:
  * under torch.distributed the gradients really are averaged over ranks (the reference wraps the
    model in DDP, train.py:1076, but calls `.module.loss`, so its reducer never fires): static
    buckets in expected-ready order are all-reduced with NCCL on a side stream as soon as the last
    gradient of a bucket has been produced, overlapping the rest of the backward pass;
  * the whole step (RNG draws, forward, backward, all-reduce, unscale, clip, AdamW, loss-scale update) is
    captured once into a CUDA graph and replayed: the GradScaler logic (found-inf check, skipped step, backoff /
    growth of the scale) runs on the device, so there are no host syncs and no per-kernel launch latency.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import kernels as K
from . import ops


def _ready_order(named_params):
    """Parameters in the order backward is expected to finish them: reverse registration order,
    with the tensors every level feeds (time MLP, relative-position table) and the input stage last."""
    late, normal = [], []
    for name, p in named_params:
        if not p.requires_grad:
            continue
        if ".time_mlp." in name or ".time_rel_pos_bias." in name or ".input_conv." in name or ".input_temp_op." in name:
            late.append((name, p))
        else:
            normal.append((name, p))
    return list(reversed(normal)) + list(reversed(late))


class GradBuckets:
    """All gradients live in one flat fp32 buffer (p.grad are views), cut into `n_buckets`
    contiguous buckets in expected-ready order.  With a process group, each bucket is all-reduced
    on `comm_stream` from the hook of its last-finished parameter."""

    def __init__(self, module: torch.nn.Module, n_buckets: int = 4, process_group=None):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or dist.is_initialized()) else 1
        order = _ready_order(module.named_parameters())
        self.params = [p for _, p in order]
        self._named = order
        self._seen = {}
        # every tensor starts on a 64-byte boundary of the flat buffers (kernels read parameters with
        # 16-byte vector loads); the padding stays zero in the gradient, parameter and moment buffers
        pad = lambda n: (n + 15) // 16 * 16
        total = sum(pad(p.numel()) for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        self.bounds: List[int] = [0]
        per_bucket = -(-total // max(1, n_buckets))
        self._bucket_of = {}
        self._pending_init: List[int] = []
        self.offsets = {}                     # id(param) -> start of its slice in the flat buffers
        # what the reference's optimizer sees: AdamW(diffusion.parameters()) (train.py:1078) -- EVERY parameter in
        # registration order, including the frozen rotary `freqs` (index 3), which holds an index but never a state
        self.registration_order = list(module.parameters())
        for p in self.params:
            n = p.numel()
            self.offsets[id(p)] = off
            p.grad = self.flat[off:off + n].view_as(p)
            b = len(self.bounds) - 1
            self._bucket_of[p.data_ptr()] = b
            if len(self._pending_init) <= b:
                self._pending_init.append(0)
            self._pending_init[b] += 1
            off += pad(n)
            if off - self.bounds[-1] >= per_bucket and off < total:
                self.bounds.append(off)
        self.bounds.append(total)
        self.n_buckets = len(self.bounds) - 1
        self._pending = list(self._pending_init)
        self.on_cuda = dev.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.world > 1 and self.on_cuda) else None
        self._hooks = []
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._owned = {p.data_ptr() for p in self.params}
        # packed fp32 scratch of the multi-tap weight gradients: (param ptr, tap key) -> (tensor, spec)
        self._scratch = {}
        self._unpack_plan = {}   # bucket -> (entry count, device descriptor table)
        ops.set_grad_sink(self)  # weight gradients are accumulated straight into the flat buffer

    # grad-sink protocol (ops._wgrad_to_param)
    def owns(self, p) -> bool:
        return p.data_ptr() in self._owned

    def scratch(self, p, key, spec) -> torch.Tensor:
        """Persistent zeroed fp32 [O, T, I] buffer the tcgen05 weight-gradient kernel accumulates into;
        `unpack_bucket` adds it into p.grad (layout `spec`) and re-zeroes it."""
        k = (p.data_ptr(), key)
        ent = self._scratch.get(k)
        if ent is None:
            O, T, I = spec[0], spec[1], spec[2]
            ent = (torch.zeros((O, T, I), dtype=torch.float32, device=p.device), spec, p,
                   self._bucket_of[p.data_ptr()])
            self._scratch[k] = ent
            self._unpack_plan.clear()
        return ent[0]

    def unpack_bucket(self, b: Optional[int]) -> None:
        """One kernel adds every packed scratch of bucket `b` (all buckets if None) into the flat gradient."""
        from . import _lib
        ents = [e for e in self._scratch.values() if b is None or e[3] == b]
        if not ents:
            return
        plan = self._unpack_plan.get(b)
        if plan is None or plan[0] != len(ents):
            arr = (_lib.PackDesc * len(ents))()
            for d, (buf, spec, p, _) in zip(arr, ents):
                d.src, d.dst = buf.data_ptr(), p.grad.data_ptr()
                d.O, d.T, d.I, d.so, d.si = spec[0], spec[1], spec[2], spec[3], spec[4]
                for i, o in enumerate(spec[5]):
                    d.tap_off[i] = int(o)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            plan = (len(ents), host.to(self.flat.device))
            self._unpack_plan[b] = plan
        _lib.call("cesm_unpack_wgrads_batched", plan[1].data_ptr(), plan[0], K._stream())

    def ready(self, p) -> None:
        if self.world > 1:
            self._on_grad(p)

    def begin_step(self):
        self.flat.zero_()
        self._pending = list(self._pending_init)
        self._seen = {}

    def _on_grad(self, p):
        # A parameter whose gradient was written directly (ops._direct) reports through ready(); torch ALSO runs
        # its post-accumulate hook (measured on 2 B200s: every direct parameter reported twice, the buckets
        # reached zero after half of their gradients and were all-reduced early -- replicas diverged).  Count
        # each parameter once per step.
        if id(p) in self._seen:
            return
        b = self._bucket_of[p.data_ptr()]
        self._seen[id(p)] = True
        self._pending[b] -= 1
        if self._pending[b] == 0:
            if self.on_cuda and ops._SIDE[0] is not None:
                torch.cuda.current_stream().wait_stream(ops._SIDE[0])   # weight gradients issued on the side stream
            self.unpack_bucket(b)
            bucket = self.flat[self.bounds[b]:self.bounds[b + 1]]
            if not self.on_cuda:  # host tensors (gloo): used by the CPU tests of the bucket logic
                dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.pg)
                return
            cur = torch.cuda.current_stream()
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.pg)

    def finish_step(self):
        if self.on_cuda:
            ops.join_side_stream()
        if self.world == 1:
            self.unpack_bucket(None)
        if self.world > 1:
            # a parameter that produced no gradient leaves its bucket un-reduced and the replicas diverge
            # silently (the reference passes find_unused_parameters=True for this, train.py:1076): fail loudly.
            # Host-side bookkeeping only; during graph capture it is checked once, for the captured step.
            late = [b for b, n in enumerate(self._pending) if n > 0]
            if late:
                names = [n for n, p in self._named if self._bucket_of.get(p.data_ptr()) in late and not self._seen.get(id(p))]
                raise RuntimeError(f"gradient buckets {late} were never all-reduced: no gradient arrived for {names[:8]}"
                                   f"{' ...' if len(names) > 8 else ''}")
        if self.world > 1 and self.on_cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def flatten_params_(self) -> torch.Tensor:
        """Re-home every parameter into one flat fp32 buffer laid out like `flat` (same offsets), so the
        optimizer step is one elementwise kernel over four flat arrays.  `p.data` become views: state
        dicts, `load_state_dict` (in-place copies) and the modules themselves see no difference."""
        assert not self._scratch, "flatten the parameters before the first step"
        flat_p = torch.zeros_like(self.flat)
        off = 0
        bucket_of = {}
        for p in self.params:
            n = p.numel()
            view = flat_p[off:off + n].view_as(p)
            view.copy_(p.data)
            b = self._bucket_of[p.data_ptr()]
            p.data = view
            bucket_of[p.data_ptr()] = b  # the sink protocol is keyed on the (new) storage address
            off += (n + 15) // 16 * 16
        self._bucket_of = bucket_of
        self._owned = set(bucket_of)
        return flat_p

    def clip_(self, max_norm: float) -> torch.Tensor:
        """Global-norm clip of every gradient (train.py:865) on the flat buffer: two kernels, no sync."""
        total = torch.linalg.vector_norm(self.flat)
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        self.flat.mul_(coef)
        return total


class FusedAdamW:
    """torch.optim.AdamW (train.py:1078-1083) + torch.amp.GradScaler (train.py:862-867, 1084) + clip_grad_norm_
    (train.py:865) as three launches of the library's own kernels over flat buffers (`cesm_adamw_step`).  Step
    counter, loss scale, growth tracker, learning rate and weight decay live in a small device vector (`state`,
    layout in include/cesm_b200.h) so that a replayed CUDA graph advances / honours them.  Drop-in for what the
    engine and train.py use of an optimizer: `step()`, `state_dict()`, `load_state_dict()`,
    `param_groups[0]["lr"]` (picked up by the next step through `sync_hparams`, also after graph capture)."""

    def __init__(self, buckets: "GradBuckets", lr, betas, weight_decay, eps, max_grad_norm,
                 init_scale: float = 65536.0, growth_interval: int = 2000):
        from . import _lib
        self._buckets = buckets
        self.g = buckets.flat
        self.p = buckets.flatten_params_()
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        st = torch.zeros(K.OPT_STATE_FLOATS, dtype=torch.float32)
        st[K.OPT_SCALE], st[K.OPT_INTERVAL] = init_scale, growth_interval  # GradScaler() defaults
        self.state = st.to(self.p.device)
        self.partials = torch.zeros(_lib.load().cesm_adamw_partials(), dtype=torch.float32, device=self.p.device)
        self.param_groups = [dict(lr=lr, betas=tuple(betas), weight_decay=weight_decay, eps=eps)]
        self.max_grad_norm = max_grad_norm
        self._synced = None    # (lr, weight_decay) last written to the device
        self._frozen = None    # (betas, eps, max_norm) baked into the first launch / the captured graph
        self.sync_hparams()

    @property
    def grad_norm(self) -> torch.Tensor:
        return self.state[K.OPT_NORM]

    @property
    def loss_scale(self) -> torch.Tensor:
        """0-dim device view of the current loss scale (multiply the loss by it before backward)."""
        return self.state[K.OPT_SCALE]

    def sync_hparams(self) -> None:
        """Bring the device copy of lr / weight decay up to date with `param_groups` (a 8-byte H2D copy, only when
        they changed; call outside graph capture -- TrainEngine does, before every replay)."""
        hp = self.param_groups[0]
        cur = (float(hp["lr"]), float(hp["weight_decay"]))
        if cur != self._synced:
            if self.p.is_cuda and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdamW: lr / weight_decay changed while a CUDA graph is being captured")
            self.state[K.OPT_LR:K.OPT_WD + 1].copy_(torch.tensor(cur, dtype=torch.float32))
            self._synced = cur
        frozen = (tuple(hp["betas"]), float(hp["eps"]), self.max_grad_norm)
        if self._frozen is not None and frozen != self._frozen:
            raise RuntimeError("FusedAdamW: betas / eps / max_grad_norm are fixed after the first step "
                               f"(were {self._frozen}, now {frozen})")

    def step(self) -> None:
        hp = self.param_groups[0]
        if not (self.p.is_cuda and torch.cuda.is_current_stream_capturing()):
            self.sync_hparams()
        self._frozen = (tuple(hp["betas"]), float(hp["eps"]), self.max_grad_norm)
        K.adamw_step(self.p, self.g, self.m, self.v, self.partials, self.state, hp["betas"][0], hp["betas"][1],
                     hp["eps"], self.max_grad_norm)
        ops.invalidate_weight_cache()  # the fp16 operand copies are stale now

    def _slices(self):
        """index -> (flat offset, numel, shape) for every TRAINABLE parameter, indexed as a torch optimizer built over
        `module.parameters()` indexes them (the reference's AdamW, train.py:1078: frozen parameters keep their slot)."""
        return {i: (self._buckets.offsets[id(p)], p.numel(), p.shape)
                for i, p in enumerate(self._buckets.registration_order) if id(p) in self._buckets.offsets}

    def state_dict(self) -> dict:
        """torch.optim.AdamW's state-dict format, so that checkpoints interchange with the reference's
        `optimizer.load_state_dict` (train.py:934-939) and with any torch AdamW over the same module."""
        step = self.state[0].detach().clone()
        state = {i: {"step": step.clone(), "exp_avg": self.m[o:o + n].view(shape).clone(),
                     "exp_avg_sq": self.v[o:o + n].view(shape).clone()}
                 for i, (o, n, shape) in self._slices().items()}
        g = self.param_groups[0]
        group = dict(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"], weight_decay=g["weight_decay"], amsgrad=False,
                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     params=list(range(len(self._buckets.registration_order))))
        # "grad_scaler": torch.amp.GradScaler.state_dict() keys (the reference loads one if present, train.py:940-944)
        scaler = {"scale": float(self.state[K.OPT_SCALE]), "growth_factor": 2.0, "backoff_factor": 0.5,
                  "growth_interval": int(self.state[K.OPT_INTERVAL]), "_growth_tracker": int(self.state[K.OPT_TRACKER])}
        return {"state": state, "param_groups": [group], "grad_scaler": scaler}

    def load_state_dict(self, sd: dict) -> None:
        if "state" in sd:  # torch format (ours, the reference's, or any torch AdamW over module.parameters())
            slices = self._slices()
            steps = []
            for i, (o, n, shape) in slices.items():
                st = sd["state"].get(i)
                if st is None:
                    continue
                self.m[o:o + n].view(shape).copy_(st["exp_avg"])
                self.v[o:o + n].view(shape).copy_(st["exp_avg_sq"])
                steps.append(float(st["step"]))
            if steps:
                self.state[0] = max(steps)
            for g, src in zip(self.param_groups, sd.get("param_groups", [])):
                g.update({k: src[k] for k in ("lr", "betas", "eps", "weight_decay") if k in src})
            sc = sd.get("grad_scaler")
            if sc:
                self.state[K.OPT_SCALE] = float(sc["scale"])
                self.state[K.OPT_TRACKER] = float(sc.get("_growth_tracker", 0))
                self.state[K.OPT_INTERVAL] = float(sc.get("growth_interval", 2000))
            self.sync_hparams()
            return
        self.state[:1].copy_(sd["step"])  # flat format of earlier checkpoints of this repo
        self.m.copy_(sd["exp_avg"])
        self.v.copy_(sd["exp_avg_sq"])
        for g, src in zip(self.param_groups, sd.get("param_groups", [])):
            g.update(src)
        self.sync_hparams()


class TrainEngine:
    """One data-parallel training step of `Diffusion.loss` as a replayable CUDA graph."""

    def __init__(self, diffusion, batch_shape, cond_shape, lr: float = 2e-4, betas=(0.9, 0.999),
                 weight_decay: float = 1e-4, eps: float = 1e-8, max_grad_norm: Optional[float] = 1.0,
                 n_buckets: int = 4, use_graph: bool = True, process_group=None):
        self.diffusion = diffusion
        dev = next(diffusion.parameters()).device
        self.device = dev
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.pg = process_group
        if self.world > 1:
            for t in list(diffusion.parameters()) + list(diffusion.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)  # what DDP's constructor does
        self.buckets = GradBuckets(diffusion, n_buckets=n_buckets, process_group=process_group)
        self.max_grad_norm = max_grad_norm
        # AdamW as train.py:1078-1083 with the clip of train.py:865 folded in; the step lives inside the graph
        if self.buckets.on_cuda:
            self.opt = FusedAdamW(self.buckets, lr, betas, weight_decay, eps, max_grad_norm)
        else:  # host tensors: only the gloo tests of the bucket logic come through here
            self.opt = torch.optim.AdamW(self.buckets.params, lr=lr, betas=betas, weight_decay=weight_decay, eps=eps)
        self.x0 = torch.zeros(batch_shape, dtype=torch.float32, device=dev)
        self.cond = torch.zeros(cond_shape, dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)
        self.use_graph = use_graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self._warm = 0
        self.arena = K.ZeroArena(dev) if dev.type == "cuda" else None
        # weight gradients on a second stream (ops._on_side); CESM_NO_SIDE_STREAM=1 keeps everything on one stream
        self.side_stream = (torch.cuda.Stream(device=dev)
                            if dev.type == "cuda" and not int(__import__("os").environ.get("CESM_NO_SIDE_STREAM", "0")) else None)

    # the captured region ----------------------------------------------------------------------
    def _step_body(self):
        if self.arena is None:
            return self._step_body_inner()
        try:
            self.arena.begin()  # one memset; the library's per-call scratch memsets are off until end()
            self._step_body_inner()
        finally:
            self.arena.end()
            K.STABLE_WEIGHT_PTRS = set()
            ops.set_side_stream(None)

    def _step_body_inner(self):
        ops.set_grad_sink(self.buckets)
        ops.set_side_stream(self.side_stream)
        ops.prepack_all()  # one kernel refreshes every fp16 operand copy of the (just updated) weights
        K.STABLE_WEIGHT_PTRS = ops.packed_ptrs()  # nothing rewrites them until the next step
        self.buckets.begin_step()
        loss = self.diffusion.loss(self.x0, self.cond)
        if isinstance(self.opt, FusedAdamW):
            # scaler.scale(loss).backward() (train.py:862); /world turns the bucket SUM into the mean
            (loss * (self.opt.loss_scale * (1.0 / self.world))).backward()
        else:
            (loss / self.world if self.world > 1 else loss).backward()
        self.buckets.finish_step()
        if isinstance(self.opt, FusedAdamW):
            self.opt.step()
            self.grad_norm.copy_(self.opt.grad_norm)
        else:
            if self.max_grad_norm is not None:
                self.grad_norm.copy_(self.buckets.clip_(self.max_grad_norm))
            self.opt.step()
        self.loss.copy_(loss.detach())

    def _run(self):
        from . import _lib
        if not self.use_graph:
            n0 = _lib.launch_count()
            self._step_body()
            self.launches_per_step = _lib.launch_count() - n0
            return
        if self.graph is None:
            if self._warm < 2:  # eager warm-up: lazy init of optimizer state, NCCL, kernel attributes
                self._warm += 1
                self._step_body()
                return
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            # (capturing the critical chain on a higher-priority stream than the weight-gradient stream was tried:
            # 11.37 ms against 10.98 ms with equal priorities -- the interleaving the scheduler picks by itself is better)
            with torch.cuda.graph(self.graph):
                self._step_body()
            self.launches_per_step = _lib.launch_count() - n0
        if isinstance(self.opt, FusedAdamW):
            self.opt.sync_hparams()  # an lr schedule / warm-up changes param_groups between replays
        self.graph.replay()
        # the replayed optimizer step rewrote the parameters behind torch's back: every cached fp16 operand
        # copy / F = 1 fold keyed on (pointer, version, epoch) is stale for any eager or eval use that follows
        ops.invalidate_weight_cache()

    # public -----------------------------------------------------------------------------------
    def step_resident(self) -> torch.Tensor:
        """One step on whatever is in the static device buffers `x0` / `cond`."""
        self._run()
        return self.loss

    def step_indices(self, device_ds, indices, augment: bool = True) -> torch.Tensor:
        """One step on samples of a device-resident dataset (`synthetic.DeviceEnsemble`): the batch is
        gathered straight into the static buffers by one kernel; only a 24-byte plan per sample crosses PCIe."""
        device_ds.batch_into(indices, self.cond, self.x0, augment)
        self._run()
        return self.loss

    def step(self, x0: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
        """One step on a host (ideally pinned) or device batch; returns the device loss scalar."""
        self.x0.copy_(x0, non_blocking=True)
        self.cond.copy_(cond, non_blocking=True)
        self._run()
        return self.loss


class SampleEngine:
    """Ensemble generation (inference.py:217-232 + model.py:186-194): the reverse chain for a batch
    of independent fields, one UNet call + fused p_sample update per step, replayed as a CUDA graph
    whose only per-step inputs are the timestep vector and the fresh noise (device RNG)."""

    def __init__(self, diffusion, shape, use_graph: bool = True):
        self.diffusion = diffusion
        dev = next(diffusion.parameters()).device
        self.device = dev
        self.shape = tuple(shape)
        self.x = torch.zeros(self.shape, dtype=torch.float32, device=dev)
        self.cond = torch.zeros(self.shape, dtype=torch.float32, device=dev)
        self.z = torch.zeros(self.shape, dtype=torch.float32, device=dev)
        self.t = torch.zeros((self.shape[0],), dtype=torch.long, device=dev)
        self.arena = K.ZeroArena(dev) if dev.type == "cuda" else None
        self.use_graph = use_graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0

    def _body(self):
        if self.arena is None:
            return self._body_inner()
        try:
            self.arena.begin()  # scratch of the step comes zeroed from one arena: one memset, not ~30
            self._body_inner()
        finally:
            self.arena.end()

    def _body_inner(self):
        d = self.diffusion
        with torch.no_grad():
            eps = d.model(self.x, self.cond, self.t)
            self.z.normal_()  # model.py:181 randn_like; kept in a static buffer so that a caller can read the draw
            self.x.copy_(K.p_sample(self.x, eps, self.z, self.t, d.betas, d.sqrt_one_minus_alphas_cumprod,
                                    d.sqrt_recip_alphas, d.posterior_variance))
            self.t.sub_(1)

    def step(self):
        from . import _lib
        if not self.use_graph:
            n0 = _lib.launch_count()
            self._body()
            self.launches_per_step = _lib.launch_count() - n0
            return
        if self.graph is None:
            x_keep, t_keep = self.x.clone(), self.t.clone()
            self._body()  # eager warm-up
            torch.cuda.synchronize()
            self.x.copy_(x_keep)
            self.t.copy_(t_keep)
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self._body()
            self.launches_per_step = _lib.launch_count() - n0
            self.x.copy_(x_keep)
            self.t.copy_(t_keep)
        self.graph.replay()

    def refresh_operands(self) -> None:
        """Operands derived from the weights (packed fp16 copies, F = 1 folds) are captured by ADDRESS: bring their
        contents up to date with whatever training / load_state_dict did since the graph was captured."""
        ops.prepack_all()
        ops.refresh_folds()

    @torch.no_grad()
    def sample(self, cond: torch.Tensor, steps: Optional[int] = None) -> torch.Tensor:
        """cond: [B,1,H,W] -> generated field [B,1,H,W] after `steps` (default T) reverse steps."""
        T = self.diffusion.T
        steps = T if steps is None else steps
        self.refresh_operands()
        self.cond.copy_(cond, non_blocking=True)
        self.x.normal_()
        self.t.fill_(T - 1)
        for _ in range(steps):
            self.step()
        return self.x.clone()
