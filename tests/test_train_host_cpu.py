"""CPU (no GPU): host logic of the training step — the dry training plan (gradient layout, backward op list), the bucket
schedule, and the world_size-2 gloo run of the bucketed gradient all-reduce + averaging the data-parallel step uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _layout(variant="small", size=64, b=2, precision="fp32"):
    from cv_diffusion_model_b200.config import variant_config
    from cv_diffusion_model_b200.training import plan_gradient_layout
    return plan_gradient_layout(variant_config(variant, size, in_channels=6), "gcd" if variant in ("tiny", "base") else "strict",
                                b, size, size, precision)


def test_dry_training_plan_covers_every_parameter():
    from cv_diffusion_model_b200.modules import create_efficient_unet
    for variant, size in (("small", 64), ("small", 256), ("tiny", 64), ("large", 128)):
        infos, total, nops = _layout(variant, size)
        m = create_efficient_unet(variant, image_size=size, in_channels=6, groupnorm="gcd" if variant == "tiny" else "strict")
        params = dict(m.named_parameters())
        assert {r[0] for r in infos} == set(params), variant
        spans = sorted((off, off + numel) for _, off, numel, _ in infos)
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= total      # disjoint, inside the buffer
        assert all(off % 4 == 0 for off, _ in spans)                                               # 16-byte aligned slices
        for name, off, numel, ready in infos:
            assert numel == params[name].numel(), name
            assert 0 <= ready < nops, (name, ready)                                                # every gradient has a producer
        # backward order: the last layers' gradients complete first
        ready = {r[0]: r[3] for r in infos}
        assert ready["final_conv.weight"] < ready["init_conv.weight"] < ready["time_mlp.1.weight"]
        assert ready["decoder_blocks.3.0.expand.weight"] < ready["encoder_blocks.0.0.expand.weight"]


def test_dry_plan_cannot_run_and_training_needs_supported_precision():
    import ctypes as C
    from cv_diffusion_model_b200 import native
    from cv_diffusion_model_b200.config import variant_config
    lib = native.lib()
    cfg = native.config_struct(variant_config("small", 64, in_channels=6), "strict")
    h = C.c_void_p()
    native.check(lib.lcm_plan_create(C.byref(cfg), 2, 64, 64, native.PREC_FP32, native.FLAG_TRAIN | native.FLAG_DRY, 0, C.byref(h)))
    with pytest.raises(ValueError, match="DRY"):
        native.check(lib.lcm_plan_set_weight(h, b"init_conv.bias", C.c_void_p(8), 32, None))
    with pytest.raises(ValueError, match="DRY"):
        native.check(lib.lcm_train_backward(h, C.c_void_p(8), 3, 0, C.c_void_p(8), 3, 0, C.c_void_p(8), C.c_void_p(8), C.c_void_p(8),
                                            0, 1.0, None, 0, -1, C.c_void_p(8), None))
    lib.lcm_plan_destroy(h)
    with pytest.raises(ValueError, match="training plans"):
        native.check(lib.lcm_plan_create(C.byref(cfg), 2, 64, 64, native.PREC_BF16,
                                         native.FLAG_TRAIN | native.FLAG_DRY | native.FLAG_SIMT_GEMM, 0, C.byref(h)))
    # an inference plan has no backward
    native.check(lib.lcm_plan_create(C.byref(cfg), 2, 64, 64, native.PREC_BF16, native.FLAG_DRY, 0, C.byref(h)))
    assert lib.lcm_train_num_backward_ops(h) == 0 and lib.lcm_train_grad_elems(h) == 0
    lib.lcm_plan_destroy(h)


def test_bucket_schedule():
    from cv_diffusion_model_b200.training import make_buckets
    infos, total, nops = _layout("small", 256, 2, "bf16")
    for n in (1, 2, 4, 8):
        buckets = make_buckets(infos, total, n)
        assert len(buckets) <= n and len(buckets) >= min(n, 2) - 1
        spans = sorted((b, e) for b, e, _ in buckets)
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))                    # a partition of the flat buffer
        assert [r for _, _, r in buckets] == sorted(r for _, _, r in buckets)        # handed over in completion order
        for b, e, ready in buckets:                                                   # complete when handed over
            assert ready == max(r[3] for r in infos if b <= r[1] < e)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cv_diffusion_model_b200.training import allreduce_buckets, make_buckets
    infos, total, _ = _layout("tiny", 64, 2, "fp32")
    buckets = make_buckets(infos, total, 4)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g)             # this rank's gradient of its shard of the global batch
    mine = flat.clone()
    works = allreduce_buckets(flat, buckets, async_op=True)      # the order the backward pass hands them over
    for w in works:
        w.wait()
    flat /= world                                       # grad_div of lcm_adamw_ema_step
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    want = sum(gathered) / world
    # the clipped global norm is the same on every rank without another collective
    norms = [None] * world
    dist.all_gather_object(norms, flat.norm().item())
    if rank == 0:
        q.put((torch.allclose(flat, want, atol=1e-6), len(buckets), max(norms) - min(norms)))
    dist.destroy_process_group()


def test_two_rank_gloo_bucketed_allreduce():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, nb, spread = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and nb >= 2 and spread == 0.0
