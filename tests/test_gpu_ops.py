"""-m gpu: single-kernel parity through the C ABI against plain torch fp32 ops on the same inputs."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _xform(x, coef, mode):
    if coef is None or mode == 0:
        return x
    if mode == 4:          # SE gate: pure scale (the b slot is defined to be 0)
        return x * coef[..., 0]
    y = x * coef[..., 0] + coef[..., 1]
    if mode == 2:
        y = y.clamp(0, 6)
    elif mode == 3:
        y = F.silu(y)
    return y


def _gemm_ref(segs, weight, P):
    cols = []
    for a, coef, mode in segs:
        M, K = a.shape
        af = a.float().view(M // P, P, K)
        cols.append(_xform(af, None if coef is None else coef[:, None], mode).reshape(M, K))
    A = torch.cat(cols, dim=1)
    return A @ weight.t()


def _stats_ref(out, P):
    o = out.double().view(out.shape[0] // P, P, out.shape[1])
    return torch.stack([o.sum(1), (o * o).sum(1)], dim=-1)


GEMM_CASES = [  # (images, P, [K...], Nc, modes)
    (2, 256, [32], 128, [2]),            # expand, level 0
    (2, 256, [128, 32], 32, [1, 0]),     # project + identity/skip
    (3, 64, [64, 32], 384, [2, 2]),      # concat expand (dec L3 b0)
    (2, 64, [384, 64, 32], 32, [1, 0, 0]),
    (1, 16, [512], 2048, [2]),           # deep level, N tiling
    (2, 16, [2048, 512], 256, [1, 0]),
    (2, 100, [48], 192, [2]),            # base-variant widths, ragged M
    (1, 16, [16], 32, [2]),              # tiny
    (2, 64, [256], 384, [1]),            # to_qkv
    (2, 1024, [64], 256, [2]),           # block_n 256 with per-image statistics kept in smem
    (3, 256, [256, 64], 64, [1, 0]),
    (2, 384, [96], 384, [2]),            # block_n 192, P multiple of 128
    (3, 256, [128, 32], 32, [4, 0]),     # SE-gated project: gate folded into the resident weights per image
    (2, 384, [384, 64, 32], 32, [4, 0, 0]),
    (2, 64, [128, 32], 64, [4, 0]),      # same, tile spans images -> A-side gating
    (2, 128, [1024, 256], 256, [4, 0]),  # weights too large to stay resident -> A-side gating
]


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("images,P,Ks,Nc,modes", GEMM_CASES)
def test_gemm(images, P, Ks, Nc, modes, dtype, impl):
    from cv_diffusion_model_b200 import ops
    if impl == 1 and os.environ.get("LCM_SKIP_TC"):
        pytest.skip("LCM_SKIP_TC set")
    if impl == 1 and dtype != torch.bfloat16:
        pytest.skip("tcgen05 kernel is bf16 only")
    g = torch.Generator(device="cuda").manual_seed(7)
    M = images * P
    segs = []
    for K, mode in zip(Ks, modes):
        a = torch.randn(M, K, device="cuda", generator=g).to(dtype)
        coef = None
        if mode != 0:
            coef = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5,
                                torch.randn(images, K, device="cuda", generator=g) * (0.0 if mode == 4 else 0.3)], dim=-1)
        segs.append((a, coef, mode))
    w = torch.randn(Nc, sum(Ks), device="cuda", generator=g) / (sum(Ks) ** 0.5)
    if dtype == torch.bfloat16:
        w = w.bfloat16().float()   # the kernels round weights to bf16; compare like with like
    out, stats = ops.gemm(segs, w, P, impl=impl)
    ref = _gemm_ref(segs, w, P)
    if dtype == torch.float32:
        assert (out - ref).abs().max().item() < 2e-4
    else:
        # bf16: the A operand is rounded to bf16 after the prologue (tensor-core input) and the output to bf16
        assert (out.float() - ref).abs().max().item() < 0.06 * ref.abs().max().item() + 0.02
        rel = ((out.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        assert rel < 8e-3, rel
    sref = _stats_ref(out.float(), P)   # statistics are defined on the stored values
    assert torch.allclose(stats, sref, rtol=1e-4, atol=1e-3), (stats - sref).abs().max()


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("N,H,W,Ci,Co", [(2, 16, 16, 32, 32), (1, 8, 12, 64, 64), (2, 8, 8, 128, 128), (1, 8, 8, 48, 48),
                                        (1, 4, 4, 256, 256), (1, 16, 32, 64, 64), (2, 32, 32, 128, 128), (3, 2, 128, 32, 32),
                                        (1, 4, 256, 96, 96), (2, 8, 128, 64, 64), (1, 6, 256, 64, 64), (2, 4, 384, 16, 16)])
def test_conv3x3(N, H, W, Ci, Co, mode, dtype, impl):
    from cv_diffusion_model_b200 import ops
    if impl == 1 and os.environ.get("LCM_SKIP_TC"):
        pytest.skip("LCM_SKIP_TC set")
    if impl == 1 and dtype != torch.bfloat16:
        pytest.skip("tcgen05 kernel is bf16 only")
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(N, H, W, Ci, device="cuda", generator=g).to(dtype)
    w = torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (9 * Ci) ** 0.5
    b = torch.randn(Co, device="cuda", generator=g) * 0.1
    if dtype == torch.bfloat16:
        w = w.bfloat16().float()
    out, stats = ops.conv3x3(x, w, b, mode, impl=impl)
    xin = x.float().permute(0, 3, 1, 2)
    if mode == 2:
        xin = F.interpolate(xin, scale_factor=2, mode="bilinear", align_corners=False)
    ref = F.conv2d(xin, w, b, stride=2 if mode == 1 else 1, padding=1).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    tol = 2e-4 if dtype == torch.float32 else 0.03
    assert (out.float() - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())
    o = out.float().double().reshape(N, -1, Co)
    sref = torch.stack([o.sum(1), (o * o).sum(1)], dim=-1)
    assert torch.allclose(stats, sref, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,H,W,C", [(2, 16, 16, 128), (1, 32, 32, 64), (2, 8, 8, 512), (1, 20, 12, 96), (3, 4, 4, 2048)])
def test_dwconv(N, H, W, C, dtype, impl):
    from cv_diffusion_model_b200 import ops
    if impl == 1 and os.environ.get("LCM_SKIP_TC"):
        pytest.skip("LCM_SKIP_TC set")
    if impl == 1 and dtype != torch.bfloat16:
        pytest.skip("tuned depthwise kernel is bf16 only")
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).to(dtype)
    coef = torch.stack([torch.rand(N, C, device="cuda", generator=g) + 0.5, torch.randn(N, C, device="cuda", generator=g)], dim=-1)
    w = torch.randn(C, 1, 3, 3, device="cuda", generator=g) / 3
    out, pool = ops.dwconv(x, coef, w, impl=impl)
    a = (x.float() * coef[:, None, None, :, 0] + coef[:, None, None, :, 1]).clamp(0, 6).permute(0, 3, 1, 2)
    ref = F.conv2d(a, w, padding=1, groups=C).permute(0, 2, 3, 1)
    tol = 1e-4 if dtype == torch.float32 else 0.05
    assert (out.float() - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())
    pref = ref.sum(dim=(1, 2))
    assert torch.allclose(pool.float(), pref, rtol=2e-3 if dtype == torch.float32 else 2e-2, atol=0.05 * H * W ** 0.5)


@pytest.mark.parametrize("N,H,W,C", [(2, 16, 16, 128), (1, 32, 32, 64), (2, 8, 8, 512), (1, 20, 12, 96), (3, 4, 4, 2048),
                                     (2, 64, 64, 128), (1, 40, 72, 256), (1, 9, 200, 128), (2, 128, 128, 384), (1, 3, 5, 32)])
def test_dwconv_f16_stream(N, H, W, C):
    """TMA-streamed fp16 depthwise kernel of the tensor-core plan (dwconv_stream.cu) vs torch fp32 conv2d."""
    from cv_diffusion_model_b200 import ops
    if os.environ.get("LCM_SKIP_TC"):
        pytest.skip("LCM_SKIP_TC set")
    g = torch.Generator(device="cuda").manual_seed(17)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).half()
    coef = torch.stack([torch.rand(N, C, device="cuda", generator=g) + 0.5, torch.randn(N, C, device="cuda", generator=g)], dim=-1)
    w = torch.randn(C, 1, 3, 3, device="cuda", generator=g) / 3
    out, pool = ops.dwconv(x, coef, w)
    assert out.dtype == torch.float16
    a = (x.float() * coef[:, None, None, :, 0] + coef[:, None, None, :, 1]).clamp(0, 6).permute(0, 3, 1, 2)
    ref = F.conv2d(a, w, padding=1, groups=C).permute(0, 2, 3, 1)
    # fp16 prologue (coefficients and result rounded to 11 bits), fp16 taps and accumulation, fp16 store
    err = (out.float() - ref).abs().max().item()
    assert err < 0.012 * max(1.0, ref.abs().max().item()), err
    rel = ((out.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert rel < 2e-3, rel
    pref = ref.sum(dim=(1, 2))
    assert torch.allclose(pool.float(), pref, rtol=5e-3, atol=0.01 * H * W ** 0.5), (pool.float() - pref).abs().max()


XDW_CASES = [  # (N, H, W, [K...], Nc)
    (2, 64, 64, [32], 128),           # level 0 block
    (3, 32, 64, [64], 256),
    (2, 64, 128, [64, 32], 384),      # decoder block on a concat input: K = 96, three n-blocks
    (4, 256, 256, [64, 32], 384),     # several items per CTA, weight reload between n-blocks, CTA ranges straddling images
    (5, 128, 128, [32, 32], 128),
    (2, 34, 64, [16], 128),           # ragged row segments (H not a multiple of the segment height)
    (1, 64, 192, [48], 256),          # base-variant widths
    (2, 64, 64, [48], 192),           # hidden width 128 + 64: the second n-block is half empty (masked stores / pool / coefficients)
    (3, 32, 128, [48, 48], 192),
]


@pytest.mark.parametrize("N,H,W,Ks,Nc", XDW_CASES)
def test_xdw_fused(N, H, W, Ks, Nc):
    """Fused expand -> GN2/ReLU6 -> depthwise path (xstats.cu + xdw_fused.cu) against the unfused kernel
    pair it replaces (gemm_expand / gemm_tc2 + dwconv_stream, themselves checked against torch above) and against torch fp32."""
    from cv_diffusion_model_b200 import ops
    if os.environ.get("LCM_SKIP_TC"):
        pytest.skip("LCM_SKIP_TC set")
    g = torch.Generator(device="cuda").manual_seed(23)
    P = H * W
    xs = [torch.randn(N, H, W, K, device="cuda", generator=g).bfloat16() for K in Ks]
    c1 = [torch.stack([torch.rand(N, K, device="cuda", generator=g) + 0.5, torch.randn(N, K, device="cuda", generator=g) * 0.5 + 1.0], -1) for K in Ks]
    Kt = sum(Ks)
    w = torch.randn(Nc, Kt, device="cuda", generator=g) / Kt ** 0.5
    c2 = torch.stack([torch.rand(N, Nc, device="cuda", generator=g) + 0.5, torch.randn(N, Nc, device="cuda", generator=g)], -1)
    wd = torch.randn(Nc, 1, 3, 3, device="cuda", generator=g) / 3
    h2, pool, stats, t = ops.xdw(list(zip(xs, c1)), w, c2, wd)
    # unfused pair
    h1, stats_ref = ops.gemm([(x.reshape(N * P, -1), c, 2) for x, c in zip(xs, c1)], w, P, impl=1, out_f16=True)
    h2_ref, pool_ref = ops.dwconv(h1.view(N, H, W, Nc), c2, wd)
    # t = relu6(a x + b) / 6 in bf16
    t_ref = torch.cat([((x.float() * c[:, None, None, :, 0] + c[:, None, None, :, 1]).clamp(0, 6) / 6) for x, c in zip(xs, c1)], -1)
    assert (t.float() - t_ref).abs().max().item() < 5e-3
    assert torch.allclose(stats, stats_ref, rtol=2e-4, atol=1e-3 * P ** 0.5), (stats - stats_ref).abs().max()
    d = (h2.float() - h2_ref.float()).abs().max().item()
    assert d == 0.0, d            # same MMAs, same fp16 roundings, same HFMA2 sequences
    assert torch.allclose(pool, pool_ref, rtol=1e-5, atol=1e-2), (pool - pool_ref).abs().max()
    # and against torch fp32
    a = (h1.float().view(N, H, W, Nc) * c2[:, None, None, :, 0] + c2[:, None, None, :, 1]).clamp(0, 6).permute(0, 3, 1, 2)
    ref = F.conv2d(a, wd, padding=1, groups=Nc).permute(0, 2, 3, 1)
    rel = ((h2.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert rel < 2e-3, rel


GEMM_F16_CASES = [  # (images, P, [K...], [segment is fp16], Nc, modes, out_f16)
    (2, 256, [32], [False], 128, [2], True),                 # expand: bf16 in, fp16 hidden out
    (3, 64, [64, 32], [False, False], 384, [2, 2], True),
    (3, 256, [128, 32], [True, False], 32, [4, 0], False),   # project: fp16 hidden (SE-gated) + bf16 residual — the level-0 shape runs on proj_stream.cu
    (7, 4096, [128, 32], [True, False], 32, [4, 0], False),  # same shape, more tiles than warps: tile rings wrap, CTA ranges straddle images
    (2, 384, [384, 64, 32], [True, False, False], 32, [4, 0, 0], False),
    (2, 64, [128, 32], [True, False], 64, [4, 0], False),    # tile spans images -> A-side gating of the fp16 operand
    (2, 128, [1024, 256], [True, False], 256, [4, 0], False),
    (2, 100, [192, 48], [True, False], 48, [4, 0], False),   # ragged M
    # streamed weights, even tile count: with LCM_PAIR=1 CTA pairs share the weight stream (multicast halves), gemm_tc2.cu `bpair`
    (64, 1024, [1024, 256], [True, False], 256, [4, 0], False),   # 32x32-level project at production size (512 tiles)
    (5, 768, [768, 192], [True, False], 128, [4, 0], False),      # 30 tiles: pairs with unequal unit counts
    (3, 256, [1024, 256], [True, False], 512, [4, 0], False),     # two n tiles
]


@pytest.mark.parametrize("images,P,Ks,h16,Nc,modes,out_f16", GEMM_F16_CASES)
def test_gemm_f16_segments(images, P, Ks, h16, Nc, modes, out_f16):
    """tcgen05 GEMM with fp16 hidden-tensor operands / outputs mixed with bf16 residual-stream operands."""
    from cv_diffusion_model_b200 import ops
    if os.environ.get("LCM_SKIP_TC"):
        pytest.skip("LCM_SKIP_TC set")
    g = torch.Generator(device="cuda").manual_seed(23)
    M = images * P
    segs, wcols = [], []
    for K, mode, is16 in zip(Ks, modes, h16):
        a = torch.randn(M, K, device="cuda", generator=g).to(torch.float16 if is16 else torch.bfloat16)
        coef = None
        if mode != 0:
            coef = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5,
                                torch.randn(images, K, device="cuda", generator=g) * (0.0 if mode == 4 else 0.3)], dim=-1)
        segs.append((a, coef, mode))
        wk = torch.randn(Nc, K, device="cuda", generator=g) / (sum(Ks) ** 0.5)
        wcols.append(wk.half().float() if is16 else wk.bfloat16().float())   # weights follow the segment's type
    w = torch.cat(wcols, dim=1)
    out, stats = ops.gemm(segs, w, P, impl=1, out_f16=out_f16)
    assert out.dtype == (torch.float16 if out_f16 else torch.bfloat16)
    ref = _gemm_ref(segs, w, P)
    assert (out.float() - ref).abs().max().item() < 0.06 * ref.abs().max().item() + 0.02
    rel = ((out.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert rel < (8e-3 if not out_f16 else 6e-3), rel
    sref = _stats_ref(out.float(), P)
    if out_f16:   # may be served by the expand kernel, whose statistics come from the input side (see test_gemm_expand_kernel)
        e0 = ((stats - sref)[..., 0].abs() / (sref[..., 1].sqrt() * P ** 0.5 + 1e-6)).max().item()
        e1 = ((stats - sref)[..., 1].abs() / sref[..., 1]).max().item()
        assert e0 < 1e-3 and e1 < 2e-3, (e0, e1)
    else:
        assert torch.allclose(stats, sref, rtol=1e-4, atol=1e-3), (stats - sref).abs().max()


EXPAND_CASES = [  # (images, P, [K...], Nc): bf16 segments, relu6 prologue, fp16 output -> gemm_expand.cu (P % 128 == 0)
    (2, 256, [32], 128),
    (3, 128, [64], 256),
    (5, 128 * 37, [32], 128),        # many tiles per CTA, image boundaries inside a CTA's range
    (2, 1024, [64, 32], 384),        # concat expand of decoder level 3 (K = 96, six n-blocks)
    (3, 640, [96], 384),
    (2, 512, [128], 512),
    (2, 384, [16], 64),
    (200, 128, [32], 128),           # more images than a CTA has tiles
    (2, 65536, [32], 128),           # production image size: 512 tiles accumulate in TMEM before the flush
    (2, 16384, [64, 32], 384),
    (3, 4096, [128], 512),           # K = 128, eight n-blocks, 128 KB of resident weights
]


WIDE_CASES = [  # the same operation on gemm_wide.cu (K a multiple of 64, 128..448; activation tile stationary)
    (2, 256, [128], 512),
    (3, 128, [256], 1024),
    (5, 128 * 9, [128, 64], 768),     # concat expand of decoder level 2
    (2, 1024, [256, 128], 1536),      # concat expand of decoder level 1
    (70, 128, [192], 256),            # more images than a CTA has tiles, three chunks
    (3, 4096, [256], 1024),
    (2, 640, [448], 128),
    (200, 128 * 3, [128], 2048),      # more tiles than CTAs: several tiles and images per CTA, 16 n-blocks
    (4, 1024, [512, 256], 3072),      # K = 768 (decoder level 0 concat): general kernel (CTA pairs sharing the weight stream with LCM_PAIR=1)
]


@pytest.mark.parametrize("images,P,Ks,Nc", EXPAND_CASES + WIDE_CASES)
def test_gemm_expand_kernel(images, P, Ks, Nc):
    """Expand GEMM specialisation: TMA store + tensor-core column statistics, x read once for all n-blocks."""
    from cv_diffusion_model_b200 import ops
    if os.environ.get("LCM_SKIP_TC"):
        pytest.skip("LCM_SKIP_TC set")
    g = torch.Generator(device="cuda").manual_seed(29)
    M = images * P
    segs = []
    for K in Ks:
        a = (torch.randn(M, K, device="cuda", generator=g) * 1.5 + 0.3).bfloat16()
        coef = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5,
                            torch.randn(images, K, device="cuda", generator=g) * 0.5 + 0.5], dim=-1)
        segs.append((a, coef, 2))
    w = (torch.randn(Nc, sum(Ks), device="cuda", generator=g) / (sum(Ks) ** 0.5)).bfloat16().float()
    out, stats = ops.gemm(segs, w, P, impl=1, out_f16=True)
    assert out.dtype == torch.float16
    ref = _gemm_ref(segs, w, P)
    assert (out.float() - ref).abs().max().item() < 0.03 * ref.abs().max().item() + 0.02
    rel = ((out.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert rel < 6e-3, rel
    # statistics: computed from the input side (column sums and Gram matrix of the prologue output) in exact
    # arithmetic they equal the sums over the fp32 accumulators; the stored values differ by fp16 rounding only
    sref = _stats_ref(out.float(), P)
    e0 = ((stats - sref)[..., 0].abs() / (sref[..., 1].sqrt() * P ** 0.5 + 1e-6)).max().item()   # |dSum| / (rms * P)
    e1 = ((stats - sref)[..., 1].abs() / sref[..., 1]).max().item()
    assert e0 < 1e-3 and e1 < 2e-3, (e0, e1)


@pytest.mark.parametrize("N,H,W", [(2, 32, 32), (1, 5, 7), (3, 33, 129), (64, 256, 256)])
def test_image_io_u8_bit_exact(N, H, W):
    """uint8 HWC <-> fp32 NCHW either side of the path (scripts/inference.py:111-127): bit-exact against the oracle,
    on the committed reference golden vectors and on seeded inputs up to the bench batch (64 x 256 x 256)."""
    import numpy as np
    from cv_diffusion_model_b200 import ops
    from oracle import image_io_oracle
    if (N, H, W) == (2, 32, 32):
        kat = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_io_kat.npz"))
        rgb, y = kat["rgb"], kat["y"]
        assert np.array_equal(ops.image_preprocess_u8(torch.from_numpy(rgb).cuda()).cpu().numpy(), kat["pre"])
        assert np.array_equal(ops.image_postprocess_u8(torch.from_numpy(y).cuda()).cpu().numpy(), kat["post"])
    rng = np.random.default_rng(N * 1000 + H)
    rgb = rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)
    y = rng.uniform(-1.2, 1.2, size=(N, 3, H, W)).astype(np.float32)
    y.reshape(-1)[:4] = [-1.0, 1.0, np.float32(2.0 / 255 - 1), -1.0000001]
    pre = ops.image_preprocess_u8(torch.from_numpy(rgb).cuda())
    assert pre.dtype == torch.float32 and tuple(pre.shape) == (N, 3, H, W)
    assert np.array_equal(pre.cpu().numpy(), image_io_oracle.preprocess_u8(rgb))
    post = ops.image_postprocess_u8(torch.from_numpy(y).cuda())
    assert post.dtype == torch.uint8 and tuple(post.shape) == (N, H, W, 3)
    assert np.array_equal(post.cpu().numpy(), image_io_oracle.postprocess_u8(y))
    with pytest.raises(ValueError):
        ops.image_preprocess_u8(torch.zeros(1, 3, 4, 4, dtype=torch.uint8, device="cuda"))


def test_image_resize_u8_bit_exact():
    """cv2.resize (INTER_LINEAR, uint8 x 3) on the device: bit-exact on the committed cv2 outputs and, against the
    oracle, on seeded up- and down-scaling cases including photo-sized inputs, a batch, 1-pixel-wide images and identity."""
    import numpy as np
    from cv_diffusion_model_b200 import ops
    from oracle import image_io_oracle
    kat = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_io_kat.npz"))
    k = 0
    while f"rs{k}_src" in kat:
        src, dst = kat[f"rs{k}_src"], kat[f"rs{k}_dst"]
        got = ops.image_resize_u8(torch.from_numpy(src[None]).cuda(), dst.shape[0], dst.shape[1])[0].cpu().numpy()
        assert np.array_equal(got, dst), k
        k += 1
    assert k >= 7
    rng = np.random.default_rng(5)
    for n, sh, sw, dh, dw in [(1, 1080, 1920, 256, 256), (3, 256, 256, 270, 480), (2, 37, 1, 64, 64), (2, 1, 9, 16, 16),
                              (4, 128, 128, 64, 64), (1, 200, 300, 200, 300), (2, 511, 513, 256, 256)]:
        src = rng.integers(0, 256, size=(n, sh, sw, 3), dtype=np.uint8)
        got = ops.image_resize_u8(torch.from_numpy(src).cuda(), dh, dw).cpu().numpy()
        assert np.array_equal(got, image_io_oracle.resize_bilinear_u8(src, dh, dw)), (n, sh, sw, dh, dw)
    ident = torch.from_numpy(rng.integers(0, 256, size=(1, 40, 50, 3), dtype=np.uint8)).cuda()
    assert torch.equal(ops.image_resize_u8(ident, 40, 50), ident)
