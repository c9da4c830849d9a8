"""Training step of the native path (BASELINE config 5: Small, 256x256, data-parallel, NCCL gradient all-reduce).

Reference semantics
  * loss          ``LowLightDiffusion.compute_loss`` / ``forward``  (src/models/low_light_diffusion.py:115-175, 250-277):
                  t ~ randint(0, T), noise ~ randn_like(high), x_t = add_noise(high, noise, t),
                  eps = unet(cat([x_t, low]), t), loss = mse | huber | l1 (eps, noise)
  * step          ``LowLightTrainer.train_epoch``  (src/training/trainer.py:283-322): zero_grad, backward,
                  clip_grad_norm_(1.0), AdamW(lr 1e-4, wd 0.01).step, EMA(0.9999).update
  * data parallel not in the reference (single device, trainer.py:142): the global batch is sharded by image, one
                  process per GPU, gradients are summed with NCCL and divided by the world size.

Two ways in:
  * ``LowLightDiffusion.compute_loss(low, high)`` returns a loss tensor whose ``backward()`` runs the native backward
    pass and fills ``param.grad`` — the reference's trainer loop (any torch optimizer, GradScaler, clip_grad_norm_) works
    unchanged on top of it.
  * ``NativeTrainer.train_step(low, high)`` keeps parameters, gradients, Adam moments and the EMA shadow in flat fp32
    buffers and runs backward -> bucketed all-reduce (overlapped with the rest of the backward pass) -> one fused
    clip + AdamW + EMA kernel -> re-pack, with no per-parameter host work.

Everything numerical is in ``liblcmunet.so`` (csrc/train_kernels.cu, csrc/plan.cu); PyTorch owns tensors, streams and
the process group.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import native
from .engine import _PREC, _require_cuda, _stream_ptr

LOSS_TYPES = {"mse": 0, "l1": 1, "huber": 2}
_UPSTREAM = 3      # lcm_train_backward: target_dev holds d loss / d eps


def plan_gradient_layout(config, groupnorm: str, batch: int, height: int, width: int, precision: str = "bf16"
                         ) -> Tuple[List[Tuple[str, int, int, int]], int, int]:
    """(name, element offset, numel, ready_after_op) of every weight gradient in the flat buffer, total elements and
    number of backward ops — from a dry plan (host only, no GPU needed)."""
    lib = native.lib()
    cfg = native.config_struct(config, groupnorm)
    h = C.c_void_p()
    native.check(lib.lcm_plan_create(C.byref(cfg), batch, height, width, _PREC[precision],
                                     native.FLAG_TRAIN | native.FLAG_DRY, 0, C.byref(h)))
    try:
        return _grad_infos(lib, h), lib.lcm_train_grad_elems(h), lib.lcm_train_num_backward_ops(h)
    finally:
        lib.lcm_plan_destroy(h)


def _grad_infos(lib, handle) -> List[Tuple[str, int, int, int]]:
    name, off, numel, ready = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int()
    out = []
    for i in range(lib.lcm_plan_num_weights(handle)):
        native.check(lib.lcm_train_grad_info(handle, i, C.byref(name), C.byref(off), C.byref(numel), C.byref(ready)))
        out.append((name.value.decode(), off.value, numel.value, ready.value))
    return out


def make_buckets(infos: Sequence[Tuple[str, int, int, int]], total_elems: int, n_buckets: int) -> List[Tuple[int, int, int]]:
    """Split the flat gradient buffer into ``n_buckets`` contiguous slices of similar size, cut at weight boundaries.
    Returns (begin, end, ready_after_op) sorted by the op index after which the slice is complete — the order in which
    the backward pass can hand them to the all-reduce."""
    infos = sorted(infos, key=lambda r: r[1])
    n_buckets = max(1, min(n_buckets, len(infos)))
    target = total_elems / n_buckets
    buckets, begin, ready = [], 0, -1
    for i, (_, off, numel, rdy) in enumerate(infos):
        ready = max(ready, rdy)
        end = infos[i + 1][1] if i + 1 < len(infos) else total_elems
        if end - begin >= target or i + 1 == len(infos):
            buckets.append((begin, end, ready))
            begin, ready = end, -1
    return sorted(buckets, key=lambda b: b[2])


def allreduce_buckets(flat: torch.Tensor, buckets: Sequence[Tuple[int, int, int]], group=None, async_op: bool = False):
    """Sum-all-reduce every bucket slice of ``flat`` (works on CPU/gloo and CUDA/NCCL tensors)."""
    import torch.distributed as dist
    works = [dist.all_reduce(flat[b:e], op=dist.ReduceOp.SUM, group=group, async_op=async_op) for b, e, _ in buckets]
    return works if async_op else None


class TrainEngine:
    """One native TRAINING plan = (weights of one EfficientUNet, batch, height, width, precision, device)."""

    def __init__(self, unet, batch: int, height: int, width: int, precision: str = "bf16", taps: bool = False,
                 device: Optional[torch.device] = None):
        if precision not in _PREC:
            raise ValueError(f"Unknown precision: {precision}")
        self.lib = native.lib()
        self.unet, self.shape, self.precision = unet, (batch, height, width), precision
        dev = torch.device(device if device is not None else "cuda")
        if dev.type != "cuda":
            raise RuntimeError("the B200 path has no CPU fallback")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        cfg = native.config_struct(unet.config, unet.groupnorm)
        flags = native.FLAG_TRAIN | (native.FLAG_TAPS if taps else 0)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            native.check(self.lib.lcm_plan_create(C.byref(cfg), batch, height, width, _PREC[precision], flags,
                                                  self.device.index, C.byref(handle)))
        self.handle = handle
        self.workspace = torch.empty(self.lib.lcm_plan_workspace_bytes(handle), dtype=torch.uint8, device=self.device)
        self.grad_elems = self.lib.lcm_train_grad_elems(handle)
        goff = self.lib.lcm_train_grad_offset_bytes(handle)
        self.flat_grad = self.workspace[goff:goff + 4 * self.grad_elems].view(torch.float32)
        self.infos = _grad_infos(self.lib, handle)
        self.num_backward_ops = self.lib.lcm_train_num_backward_ops(handle)
        self._loss_buf = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._params = None
        self._weight_version = None
        self.flat_param: Optional[torch.Tensor] = None
        self.upload_weights()

    # ---- weights --------------------------------------------------------------------------------------
    def _named_params(self) -> Dict[str, torch.nn.Parameter]:
        if self._params is None:
            self._params = dict(self.unet.named_parameters())
        return self._params

    def _version(self):
        v = 0
        for p in self._named_params().values():
            v += p._version
        return (getattr(self.unet, "_weights_epoch", 0), v)

    def upload_weights(self) -> None:
        params = self._named_params()
        with torch.cuda.device(self.device):
            if self.flat_param is not None:
                native.check(self.lib.lcm_plan_set_weights_flat(self.handle, C.c_void_p(self.flat_param.data_ptr()), _stream_ptr()))
            else:
                for name, _, numel, _ in self.infos:
                    if name not in params:
                        raise ValueError(f"the model has no parameter '{name}'")
                    w = params[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
                    if w.numel() != numel:
                        raise ValueError(f"parameter '{name}' has {w.numel()} elements, the plan expects {numel}")
                    native.check(self.lib.lcm_plan_set_weight(self.handle, name.encode(), C.c_void_p(w.data_ptr()), numel,
                                                              _stream_ptr()))
                torch.cuda.current_stream().synchronize()     # the temporaries above die here
        extra = set(params) - {r[0] for r in self.infos}
        if extra:
            raise ValueError(f"parameters unknown to the native plan: {sorted(extra)[:4]}")
        self._weight_version = self._version()

    def refresh(self) -> None:
        if self._weight_version != self._version():
            self.upload_weights()

    def flatten_parameters(self) -> torch.Tensor:
        """Move every parameter into ONE flat fp32 buffer laid out like the gradient buffer (``param.data`` becomes a view
        of it).  The fused optimizer then updates the model in place and ``lcm_plan_set_weights_flat`` re-packs from it."""
        if self.flat_param is None:
            flat = torch.zeros(self.grad_elems, dtype=torch.float32, device=self.device)
            params = self._named_params()
            with torch.no_grad():
                for name, off, numel, _ in self.infos:
                    p = params[name]
                    flat[off:off + numel].copy_(p.detach().reshape(-1))
                    p.data = flat[off:off + numel].view(p.shape)
            self.flat_param = flat
        return self.flat_param

    def grads(self) -> Dict[str, torch.Tensor]:
        """Views into the flat gradient buffer, shaped like the parameters (valid until the next backward)."""
        params = self._named_params()
        return {name: self.flat_grad[off:off + numel].view(params[name].shape) for name, off, numel, _ in self.infos}

    # ---- the three calls of a step -------------------------------------------------------------------------
    def _check(self, t_: torch.Tensor, what: str, c: int):
        b, h, w = self.shape
        _require_cuda(t_, what)
        if tuple(t_.shape) != (b, c, h, w) or not t_.is_contiguous():
            raise ValueError(f"{what} must be a contiguous [{b},{c},{h},{w}] tensor, got {tuple(t_.shape)}")

    def forward(self, noisy: torch.Tensor, low: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """eps = unet(cat([noisy, low], 1), t); every activation stays in the workspace for backward()."""
        ca = noisy.shape[1]
        self._check(noisy, "noisy", ca)
        self._check(low, "low_light", self.unet.config.in_channels - ca)
        self.refresh()
        b, h, w = self.shape
        t = t.to(device=noisy.device, dtype=torch.long).contiguous()
        eps = torch.empty(b, self.unet.config.out_channels, h, w, dtype=torch.float32, device=noisy.device)
        with torch.cuda.device(self.device):
            native.check(self.lib.lcm_unet_forward(self.handle, C.c_void_p(noisy.data_ptr()), ca, noisy.stride(0),
                                                   C.c_void_p(low.data_ptr()), low.shape[1], low.stride(0),
                                                   C.c_void_p(t.data_ptr()), C.c_void_p(eps.data_ptr()),
                                                   C.c_void_p(self.workspace.data_ptr()), _stream_ptr()))
        return eps

    def loss(self, eps: torch.Tensor, target: torch.Tensor, loss_type: str = "mse") -> torch.Tensor:
        if loss_type not in LOSS_TYPES:
            raise ValueError(f"Unknown loss type: {loss_type}")
        with torch.cuda.device(self.device):
            native.check(self.lib.lcm_train_loss(C.c_void_p(eps.data_ptr()), C.c_void_p(target.data_ptr()), eps.numel(),
                                                 LOSS_TYPES[loss_type], C.c_void_p(self._loss_buf.data_ptr()), _stream_ptr()))
        return self._loss_buf[0].to(torch.float32)

    def backward(self, noisy: torch.Tensor, low: torch.Tensor, t: torch.Tensor, eps: torch.Tensor, target: torch.Tensor,
                 loss_type: str = "mse", grad_scale: float = 1.0, grad_scale_dev: Optional[torch.Tensor] = None,
                 op_begin: int = 0, op_end: int = -1) -> None:
        """Backward ops [op_begin, op_end) of the last forward(); gradients land in ``flat_grad``."""
        t = t.to(device=noisy.device, dtype=torch.long).contiguous()
        if grad_scale_dev is not None:
            grad_scale_dev = grad_scale_dev.to(device=noisy.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.device):
            native.check(self.lib.lcm_train_backward(
                self.handle, C.c_void_p(noisy.data_ptr()), noisy.shape[1], noisy.stride(0), C.c_void_p(low.data_ptr()),
                low.shape[1], low.stride(0), C.c_void_p(t.data_ptr()), C.c_void_p(eps.data_ptr()),
                C.c_void_p(target.data_ptr()), loss_type if isinstance(loss_type, int) else LOSS_TYPES[loss_type], float(grad_scale),
                C.c_void_p(grad_scale_dev.data_ptr()) if grad_scale_dev is not None else None, op_begin, op_end,
                C.c_void_p(self.workspace.data_ptr()), _stream_ptr()))

    def read_grad_tap(self, name: str, channels: int, height: int, width: int) -> torch.Tensor:
        out = torch.empty(self.shape[0], channels, height, width, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            native.check(self.lib.lcm_train_read_grad_tap(self.handle, name.encode(), C.c_void_p(out.data_ptr()),
                                                          C.c_void_p(self.workspace.data_ptr()), _stream_ptr()))
        return out

    def backward_ops(self) -> List[Tuple[str, str]]:
        name, kern = C.c_char_p(), C.c_char_p()
        out = []
        for i in range(self.num_backward_ops):
            native.check(self.lib.lcm_train_backward_op_info(self.handle, i, C.byref(name), C.byref(kern)))
            out.append((name.value.decode(), kern.value.decode()))
        return out

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.lcm_plan_destroy(self.handle)
            self.handle = None
        self.workspace = None
        self.flat_grad = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def get_train_engine(unet, batch: int, height: int, width: int, device, precision: Optional[str] = None) -> TrainEngine:
    precision = precision or unet.precision
    dev = torch.device(device)
    key = ("train", batch, height, width, precision, dev.index if dev.index is not None else torch.cuda.current_device())
    cache = unet._engines
    eng = cache.get(key)
    if eng is None:
        from .engine import MAX_ENGINES_PER_UNET
        while len(cache) >= max(1, MAX_ENGINES_PER_UNET):
            _, old = cache.popitem(last=False)
            old.close()
        eng = TrainEngine(unet, batch, height, width, precision, device=dev)
        cache[key] = eng
    else:
        cache.move_to_end(key)
    return eng


class _NativeLoss(torch.autograd.Function):
    """loss = L(unet(cat([noisy, low]), t), target) with the native forward and backward; differentiable w.r.t. the UNet
    parameters (the inputs need no gradient in the reference's training loop)."""

    @staticmethod
    def forward(ctx, engine: TrainEngine, noisy, low, t, target, loss_type, names, *params):
        eps = engine.forward(noisy, low, t)
        loss = engine.loss(eps, target, loss_type)
        ctx.engine, ctx.loss_type, ctx.names = engine, loss_type, names
        ctx.saved = (noisy, low, t, eps, target)
        ctx.mark_non_differentiable(eps)
        return loss, eps

    @staticmethod
    def backward(ctx, grad_loss, _grad_eps):
        eng = ctx.engine
        noisy, low, t, eps, target = ctx.saved
        eng.backward(noisy, low, t, eps, target, ctx.loss_type, 1.0, grad_loss.reshape(1))
        g = eng.grads()
        # clones: autograd keeps what it is handed as .grad, the flat buffer is reused by the next backward pass
        return (None,) * 7 + tuple(g[n].clone() for n in ctx.names)


class _NativeUNet(torch.autograd.Function):
    """eps = unet(cat([xa, xb]), t), differentiable w.r.t. the UNet parameters for ANY downstream loss: backward hands the
    upstream gradient d loss / d eps to the native backward pass (loss_type 3)."""

    @staticmethod
    def forward(ctx, engine: TrainEngine, xa, xb, t, names, *params):
        eps = engine.forward(xa, xb, t)
        ctx.engine, ctx.names = engine, names
        ctx.saved = (xa, xb, t, eps)
        return eps

    @staticmethod
    def backward(ctx, grad_eps):
        eng = ctx.engine
        xa, xb, t, eps = ctx.saved
        eng.backward(xa, xb, t, eps, grad_eps.to(torch.float32).contiguous(), _UPSTREAM)
        g = eng.grads()
        return (None,) * 5 + tuple(g[n].clone() for n in ctx.names)


def native_unet_forward(unet, x: torch.Tensor, t: torch.Tensor, precision: Optional[str] = None) -> torch.Tensor:
    """``unet(x, t)`` with autograd (training mode): x = cat([xa, xb]) is split at out_channels like the pipeline's concat."""
    _require_cuda(x, "x")
    b, cin, h, w = x.shape
    ca = unet.config.out_channels if cin > unet.config.out_channels else cin
    xa = x[:, :ca].contiguous()
    xb = x[:, ca:].contiguous() if cin > ca else x.new_zeros(b, 0, h, w)
    if xb.shape[1] == 0:
        raise ValueError("training plans take the conditioning concat as two tensors (in_channels > out_channels)")
    eng = get_train_engine(unet, b, h, w, x.device, precision)
    named = list(unet.named_parameters())
    return _NativeUNet.apply(eng, xa, xb, t, tuple(n for n, _ in named), *[p for _, p in named])


def native_loss(unet, noisy: torch.Tensor, low: torch.Tensor, t: torch.Tensor, target: torch.Tensor, loss_type: str = "mse",
                precision: Optional[str] = None):
    """(loss, eps) with autograd wired to the native backward pass."""
    if loss_type not in LOSS_TYPES:
        raise ValueError(f"Unknown loss type: {loss_type}")
    _require_cuda(noisy, "noisy")
    b, _, h, w = noisy.shape
    eng = get_train_engine(unet, b, h, w, noisy.device, precision)
    named = [(n, p) for n, p in unet.named_parameters()]
    names = tuple(n for n, _ in named)
    return _NativeLoss.apply(eng, noisy.contiguous(), low.to(torch.float32).contiguous(), t, target.contiguous(), loss_type,
                             names, *[p for _, p in named])


class NativeTrainer:
    """The inner loop of ``LowLightTrainer.train_epoch`` (trainer.py:283-322) on flat buffers.

    ``train_step(low, high)``: loss forward -> native backward in chunks, each finished gradient bucket handed to an
    asynchronous all-reduce on a side stream -> global-norm clip + AdamW + EMA in one kernel -> weights re-packed.
    Defaults are the reference's ``TrainingConfig`` (lr 1e-4, weight_decay 0.01, gradient_clip 1.0, ema_decay 0.9999)
    and ``torch.optim.AdamW``'s (betas (0.9, 0.999), eps 1e-8).
    """

    def __init__(self, model, batch: int, lr: float = 1e-4, weight_decay: float = 0.01, betas=(0.9, 0.999), eps: float = 1e-8,
                 gradient_clip: float = 1.0, ema_decay: Optional[float] = 0.9999, loss_type: str = "mse",
                 precision: Optional[str] = None, process_group=None, n_buckets: int = 4, device=None):
        import torch.distributed as dist
        if loss_type not in LOSS_TYPES:
            raise ValueError(f"Unknown loss type: {loss_type}")
        self.model, self.loss_type = model, loss_type
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.gradient_clip, self.ema_decay = gradient_clip, ema_decay
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        s = model.image_size
        self.engine = TrainEngine(model.unet, batch, s, s, precision or model.unet.precision, device=dev)
        self.flat = self.engine.flatten_parameters()
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.ema = self.flat.clone() if ema_decay is not None else None
        self.step_count = 0
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.buckets = make_buckets(self.engine.infos, self.engine.grad_elems, n_buckets if self.world > 1 else 1)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=self.engine.device)
        self._comm = torch.cuda.Stream(device=self.engine.device) if self.world > 1 else None

    # EMAModel.apply_shadow / restore equivalents (trainer.py:106-118) on the flat buffers
    def ema_state(self) -> Dict[str, torch.Tensor]:
        params = dict(self.model.unet.named_parameters())
        return {n: self.ema[o:o + k].view(params[n].shape) for n, o, k, _ in self.engine.infos} if self.ema is not None else {}

    def train_step(self, low: torch.Tensor, high: torch.Tensor, timesteps: Optional[torch.Tensor] = None,
                   noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        eng, lib = self.engine, self.engine.lib
        sched = self.model.scheduler
        b = low.shape[0]
        if timesteps is None:
            timesteps = torch.randint(0, sched.config.num_train_timesteps, (b,), device=low.device)     # :143-146
        if noise is None:
            noise = torch.randn_like(high)                                                              # :150
        noisy = sched.add_noise(high.contiguous(), noise.contiguous(), timesteps)                       # :153
        low = low.to(torch.float32).contiguous()
        eps = eng.forward(noisy, low, timesteps)
        loss = eng.loss(eps, noise, self.loss_type)
        # ---- backward, gradient buckets all-reduced as they complete -------------------------------------
        if self.world > 1:
            import torch.distributed as dist
            cur = torch.cuda.current_stream(eng.device)
            begin = 0
            for (lo, hi, ready) in self.buckets:
                end = min(max(ready + 1, begin), eng.num_backward_ops)
                if end > begin:
                    eng.backward(noisy, low, timesteps, eps, noise, self.loss_type, op_begin=begin, op_end=end)
                    begin = end
                ev = torch.cuda.Event()
                ev.record(cur)
                with torch.cuda.stream(self._comm):
                    self._comm.wait_event(ev)
                    dist.all_reduce(eng.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
            if begin < eng.num_backward_ops:
                eng.backward(noisy, low, timesteps, eps, noise, self.loss_type, op_begin=begin, op_end=-1)
            cur.wait_stream(self._comm)
        else:
            eng.backward(noisy, low, timesteps, eps, noise, self.loss_type)
        # ---- clip_grad_norm_ + AdamW + EMA (:296-302, :321-322) --------------------------------------------
        self.step_count += 1
        with torch.cuda.device(eng.device):
            st = _stream_ptr()
            native.check(lib.lcm_grad_sumsq(C.c_void_p(eng.flat_grad.data_ptr()), eng.grad_elems, C.c_void_p(self._sumsq.data_ptr()), st))
            native.check(lib.lcm_adamw_ema_step(
                C.c_void_p(self.flat.data_ptr()), C.c_void_p(eng.flat_grad.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
                C.c_void_p(self.exp_avg_sq.data_ptr()), C.c_void_p(self.ema.data_ptr()) if self.ema is not None else None,
                eng.grad_elems, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
                self.ema_decay if self.ema_decay is not None else 0.0, C.c_void_p(self._sumsq.data_ptr()), float(self.world),
                self.gradient_clip if self.gradient_clip else 0.0, st))
        eng.upload_weights()                      # re-pack from the flat buffer (forward + transposed images)
        self.model.unet.mark_weights_changed()    # inference plans of the same model re-pack on their next call
        eng._weight_version = eng._version()
        return loss

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last step (after the all-reduce, before clipping), as clip_grad_norm_ returns it."""
        return (self._sumsq.sqrt() / self.world).to(torch.float32)
