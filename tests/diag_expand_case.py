import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import ops
images, P, Ks, Nc = int(sys.argv[1]), int(sys.argv[2]), [int(k) for k in sys.argv[3].split(',')], int(sys.argv[4])
g = torch.Generator(device="cuda").manual_seed(29)
M = images * P
segs = []
for K in Ks:
    a = (torch.randn(M, K, device="cuda", generator=g) * 1.5 + 0.3).bfloat16()
    coef = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5, torch.randn(images, K, device="cuda", generator=g) * 0.5 + 0.5], dim=-1)
    segs.append((a, coef, 2))
w = (torch.randn(Nc, sum(Ks), device="cuda", generator=g) / (sum(Ks) ** 0.5)).bfloat16().float()
out, stats = ops.gemm(segs, w, P, impl=1, out_f16=True)
cols = [(s[0].float().view(images, P, -1) * s[1][:, None, :, 0] + s[1][:, None, :, 1]).clamp(0, 6).reshape(M, -1) for s in segs]
ref = torch.cat(cols, 1) @ w.t()
print("out err", (out.float() - ref).abs().max().item(), "ref max", ref.abs().max().item())
o = out.float().double().view(images, P, Nc)
sref = torch.stack([o.sum(1), (o * o).sum(1)], -1)
e0 = ((stats - sref)[..., 0].abs() / (sref[..., 1].sqrt() * P ** 0.5 + 1e-6))
e1 = ((stats - sref)[..., 1].abs() / sref[..., 1])
print("e0 max", e0.max().item(), "e1 max", e1.max().item())
bad = (e1.max(dim=1).values > 2e-3).nonzero().flatten().tolist()
print("bad images", bad[:40], "of", images)
if bad:
    i = bad[0]; print(stats[i, :4], sref[i, :4])
