"""Reference-side binding of the C ABI (include/lcm_unet.h): what a maintainer of the reference would add.

Self-contained on purpose — ``ctypes`` + ``torch`` only, nothing else from this package — so the file can be copied
to ``src/models/b200_backend.py`` of zamazincode/cv-diffusion-model unchanged.  ``attach(model, batch, size)`` takes an
*unmodified* reference ``LowLightDiffusion`` (src/models/low_light_diffusion.py:31) that lives on a B200, builds a
native plan from ``model.unet.config`` + ``model.unet.state_dict()`` and replaces the execution of

    model.enhance        (low_light_diffusion.py:177-248)  ->  lcm_enhance
    model.unet.forward   (efficient_unet.py:532-606)       ->  lcm_unet_forward

keeping their signatures.  Everything else of the model (scheduler object, parameters, ``state_dict``) is untouched.
INTEGRATION.md documents the mapping; tests/test_reference_backend.py exercises this file against the unmodified
reference (build container) and against a reference-shaped model on the GPU.
"""
import ctypes as C
import os

import torch

LIB_PATH = os.environ.get("LCM_UNET_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblcmunet.so"))
PREC_FP32, PREC_BF16 = 0, 1


class UNetConfigC(C.Structure):
    """lcm_unet_config  <->  EfficientUNetConfig (efficient_unet.py:24-57)."""
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("base_channels", C.c_int32),
                ("num_levels", C.c_int32), ("channel_multipliers", C.c_int32 * 8),
                ("num_attention_resolutions", C.c_int32), ("attention_resolutions", C.c_int32 * 8),
                ("num_attention_heads", C.c_int32), ("num_res_blocks", C.c_int32), ("expansion_ratio", C.c_int32),
                ("se_ratio", C.c_float), ("time_embed_dim", C.c_int32), ("image_size", C.c_int32),
                ("groupnorm_gcd", C.c_int32), ("standard_attention", C.c_int32)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing (build it with nvcc for sm_100a); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.lcm_last_error.restype = C.c_char_p
        lib.lcm_plan_workspace_bytes.restype = C.c_size_t
        lib.lcm_plan_workspace_bytes.argtypes = [C.c_void_p]
        lib.lcm_plan_create.argtypes = [C.POINTER(UNetConfigC), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                        C.POINTER(C.c_void_p)]
        lib.lcm_plan_destroy.argtypes = [C.c_void_p]
        lib.lcm_plan_destroy.restype = None
        lib.lcm_plan_num_weights.argtypes = [C.c_void_p]
        lib.lcm_plan_weight_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64)]
        lib.lcm_plan_set_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]
        lib.lcm_unet_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.lcm_enhance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64),
                                    C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def _check(rc):
    if rc < 0:
        msg = (_load().lcm_last_error() or b"").decode()
        raise (ValueError if rc in (-1, -3) else RuntimeError)(msg)   # reference: ValueError for bad configs
    return rc


def config_struct(cfg) -> UNetConfigC:
    """Copy the reference dataclass field by field; refuses what the native path does not implement."""
    if not getattr(cfg, "use_se", True) or not getattr(cfg, "quantization_friendly", True) or getattr(cfg, "dropout", 0.0) != 0.0:
        raise ValueError("the B200 path implements the preset blocks (SE, ReLU6, dropout 0)")
    c = UNetConfigC()
    c.in_channels, c.out_channels, c.base_channels = cfg.in_channels, cfg.out_channels, cfg.base_channels
    mult, att = tuple(cfg.channel_multipliers), tuple(cfg.attention_resolutions)
    if len(mult) > 8 or len(att) > 8:
        raise ValueError("too many levels / attention resolutions")
    c.num_levels = len(mult)
    for i, m in enumerate(mult):
        c.channel_multipliers[i] = m
    c.num_attention_resolutions = len(att)
    for i, r in enumerate(att):
        c.attention_resolutions[i] = r
    c.num_attention_heads, c.num_res_blocks = cfg.num_attention_heads, cfg.num_res_blocks
    c.expansion_ratio, c.se_ratio = cfg.expansion_ratio, cfg.se_ratio
    c.time_embed_dim, c.image_size = cfg.time_embed_dim, cfg.image_size
    # GroupNorm(min(32, C), C) wherever the reference constructs; an instance of tiny/base can only exist with
    # GroupNorm(gcd(32, C), C) patched in.  gcd == min(32, C) whenever the latter divides C, so one flag
    # value serves every model instance that can be passed in.
    c.groupnorm_gcd = 1
    c.standard_attention = 0 if getattr(cfg, "use_linear_attention", True) else 1
    return c


def step_coefficients(scheduler, timesteps):
    """fp32 scalars of LCMScheduler.step for every loop iteration (lcm_scheduler.py:205-217,239-242), computed with the
    reference's own 0-d tensor arithmetic: (sqrt(1-abar_t), sqrt(abar_t), sqrt(abar_prev), sqrt(1-abar_prev))."""
    ab, coef = scheduler.alphas_cumprod, []
    n = len(timesteps)
    for i, t in enumerate(timesteps):
        prev = timesteps[i + 1] if i + 1 < n else 0
        a_t = ab[t]
        a_p = ab[prev] if prev > 0 else scheduler.final_alpha_cumprod
        coef += [float((1 - a_t) ** 0.5), float(a_t ** 0.5), float(a_p ** 0.5), float((1 - a_p) ** 0.5)]
    return coef


class Backend:
    """Owns the native plan of one (model, batch, size, precision); freed by close() / garbage collection."""

    def __init__(self, model, batch, size, precision=PREC_BF16, device=None):
        self.lib = _load()
        self.model, self.batch, self.size = model, batch, size
        dev = torch.device(device if device is not None else "cuda")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.cfg = config_struct(model.unet.config)
        self.plan = C.c_void_p()
        with torch.cuda.device(self.device):
            _check(self.lib.lcm_plan_create(C.byref(self.cfg), batch, size, size, precision, 0, self.device.index,
                                            C.byref(self.plan)))
            self.ws = torch.empty(self.lib.lcm_plan_workspace_bytes(self.plan), dtype=torch.uint8, device=self.device)
        self.upload()

    def upload(self):
        """Re-pack the weights from model.unet.state_dict() (call again after an optimizer step / EMA swap)."""
        sd = self.model.unet.state_dict()
        name, numel = C.c_char_p(), C.c_int64()
        with torch.cuda.device(self.device):
            for i in range(self.lib.lcm_plan_num_weights(self.plan)):
                _check(self.lib.lcm_plan_weight_info(self.plan, i, C.byref(name), C.byref(numel)))
                key = name.value.decode()
                if key not in sd:
                    raise ValueError(f"state_dict has no entry '{key}'")
                w = sd[key].detach().to(device=self.device, dtype=torch.float32).contiguous()
                _check(self.lib.lcm_plan_set_weight(self.plan, key.encode(), C.c_void_p(w.data_ptr()), w.numel(),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.current_stream().synchronize()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def unet_forward(self, x, timestep, return_features=False):
        if return_features:
            raise NotImplementedError("return_features is not available on the native path")
        b, s = self.batch, self.size
        if not x.is_cuda or tuple(x.shape) != (b, self.cfg.in_channels, s, s):
            raise ValueError(f"x must be a CUDA tensor of shape {(b, self.cfg.in_channels, s, s)}")
        x = x.float().contiguous()
        t = timestep.to(device=x.device, dtype=torch.long).contiguous()
        eps = torch.empty(b, self.cfg.out_channels, s, s, dtype=torch.float32, device=x.device)
        with torch.cuda.device(self.device):
            _check(self.lib.lcm_unet_forward(self.plan, C.c_void_p(x.data_ptr()), x.shape[1], x.stride(0), None, 0, 0,
                                             C.c_void_p(t.data_ptr()), C.c_void_p(eps.data_ptr()),
                                             C.c_void_p(self.ws.data_ptr()), self._stream()))
        return eps

    @torch.no_grad()
    def enhance(self, low_light, num_inference_steps=None, generator=None, return_intermediate=False):
        model, b, s = self.model, self.batch, self.size
        if not low_light.is_cuda or tuple(low_light.shape) != (b, 3, s, s):
            raise ValueError(f"low_light must be a CUDA tensor of shape {(b, 3, s, s)}")
        n = num_inference_steps or model.num_inference_steps
        model.scheduler.set_timesteps(n, device=low_light.device)                 # lcm_scheduler.py:131-167 unchanged
        ts = [int(t) for t in model.scheduler.timesteps.cpu()]
        lat = torch.randn(b, 3, s, s, device=low_light.device, generator=generator)   # low_light_diffusion.py:208-211
        noises = torch.stack([torch.randn_like(lat) for _ in range(n - 1)]) if n > 1 else None   # lcm_scheduler.py:237
        coef = step_coefficients(model.scheduler, ts)
        out = torch.empty_like(lat)
        trace = torch.empty(n, b, 3, s, s, device=lat.device) if return_intermediate else None
        low = low_light.float().contiguous()
        with torch.cuda.device(self.device):
            _check(self.lib.lcm_enhance(self.plan, C.c_void_p(low.data_ptr()), C.c_void_p(lat.data_ptr()),
                                        C.c_void_p(noises.data_ptr()) if n > 1 else None, n, (C.c_int64 * n)(*ts),
                                        (C.c_float * (4 * n))(*coef), C.c_void_p(out.data_ptr()),
                                        C.c_void_p(trace.data_ptr()) if return_intermediate else None,
                                        C.c_void_p(self.ws.data_ptr()), self._stream()))
        if return_intermediate:
            # the reference returns LowLightDiffusionOutput(enhanced=..., intermediate=[...]); mirror it when present
            cls = getattr(model, "_output_cls", None)
            inter = [trace[i] for i in range(n)]
            return cls(enhanced=out, intermediate=inter) if cls else (out, inter)
        return out

    def close(self):
        if self.plan:
            self.lib.lcm_plan_destroy(self.plan)
            self.plan = C.c_void_p()
        self.ws = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def attach(model, batch, size=None, precision=PREC_BF16, device=None):
    """Monkey-bind the native plan onto a reference ``LowLightDiffusion`` instance (on a B200) and return it."""
    size = size or model.image_size
    if getattr(model, "condition_mode", "concat") != "concat":
        raise ValueError('only condition_mode="concat" is implemented natively')
    be = Backend(model, batch, size, precision, device)
    model._b200 = be
    model.enhance = be.enhance                 # replaces low_light_diffusion.py:177-248
    model.unet.forward = be.unet_forward       # replaces efficient_unet.py:532-606
    return model
