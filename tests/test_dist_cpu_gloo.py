"""CPU, gloo, world_size=2: the data-parallel algebra the host side relies on.

Each rank evaluates ITS shard (marf_b200.engine.shard_plan) with the CPU oracle, all-reduces the 6 loss sums,
forms the per-rank surrogate whose gradient is the rank's share of d(loss.all)/dθ given the GLOBAL normalisers —
exactly what marf_step (static normalisers) and marf_step_forward/backward (implicit mask) compute per rank —
all-reduces the gradients, and the result must equal the single-process oracle step."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import fixtures as fx
    import planar_oracle as po
    from marf_b200.engine import shard_plan
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        cfg = po.PlanarConfig(**case)
        B, h, w = cfg.batch_size, cfg.h, cfg.w
        ws, bs = fx.synth_mlp(3, po.layer_shapes(cfg), scale=2.0)
        params = po.PlanarParams([t.double() for t in ws], [t.double() for t in bs], fx.synth_warp(4, B, 0.05).double())
        if cfg.use_implicit_mask:
            mw, mb = fx.synth_mlp(5, [(64, po.MASK_IN), (64, 64), (1, 64)])
            params.mask_w, params.mask_b = [t.double() for t in mw], [t.double() for t in mb]
            params.embed = fx.synth_embed(6, 8, 128).double()
        rgb, masks = fx.synth_patches(9, B, h, w, occluders=True)
        images = dict(rgb=rgb.double(), masks=masks.double() if cfg.use_masks else None, masks_eroded=None, edges=None)
        it, progress = 50, 0.25
        c_rgb, c_mask, c_edge = po.loss_coefficients(cfg, it)
        # ---------------- this rank's shard
        nb, poff, rows, roff = shard_plan(B, h, rank, world)
        leaves = params.leaves()
        for t in leaves:
            t.requires_grad_(True)
        xy = po.normalized_pixel_grid(cfg).double().view(h, w, 2)[roff:roff + rows].reshape(-1, 2)
        uv = po.warp_grid(xy.repeat(nb, 1, 1), params.warp[poff:poff + nb])
        pred = po.neural_image(uv, params.mlp_w, params.mlp_b, cfg, progress)            # [nb, rows*w, 3]
        tgt = images["rgb"][poff:poff + nb, :, roff:roff + rows].permute(0, 2, 3, 1).reshape(nb, rows * w, 3)
        if cfg.use_implicit_mask:
            m = torch.stack([po.mask_head(po.mask_features(images["rgb"][poff + i][:, roff:roff + rows].float(), xy.float(),
                                                           params.embed.float()).double(), params.mask_w, params.mask_b)
                             for i in range(nb)])                                          # [nb, rows*w, 1]
        elif cfg.use_masks:
            m = images["masks"][poff:poff + nb, :, roff:roff + rows].permute(0, 2, 3, 1).reshape(nb, rows * w, 1)
        else:
            m = torch.ones(nb, rows * w, 1, dtype=torch.float64)
        S_r = (((pred - tgt) * m) ** 2).sum()
        N_r = 3 * m.sum()
        Sm_r = ((1 - m) ** 2).sum() if cfg.use_implicit_mask else torch.zeros((), dtype=torch.float64)
        Nm_r = torch.tensor(float(nb * rows * w), dtype=torch.float64)
        sums = torch.stack([S_r.detach(), N_r.detach(), Sm_r.detach(), Nm_r])
        dist.all_reduce(sums)                                                              # the 8-double exchange
        S, N, Sm, Nm = sums
        # surrogate: d/dθ [S/N] restricted to this rank's pixels with S, N global constants
        obj = c_rgb * (S_r / N - (S / N ** 2) * N_r)
        if cfg.use_implicit_mask:
            obj = obj + c_mask * Sm_r / Nm
        obj.backward()
        flat = torch.cat([(t.grad if t.grad is not None else torch.zeros_like(t)).reshape(-1) for t in leaves])
        dist.all_reduce(flat)                                                              # the gradient exchange
        if rank == 0:
            # single-process truth
            p1 = po.PlanarParams([t.detach().clone() for t in params.mlp_w], [t.detach().clone() for t in params.mlp_b],
                                 params.warp.detach().clone())
            if cfg.use_implicit_mask:
                p1.mask_w = [t.detach().clone() for t in params.mask_w]
                p1.mask_b = [t.detach().clone() for t in params.mask_b]
                p1.embed = params.embed.detach().clone()
            _, loss, grads = po.step(p1, images, cfg, it=it, progress=progress)
            ref = torch.cat([g.reshape(-1) for g in grads])
            ret["gerr"] = float((flat - ref).abs().max() / ref.abs().max())
            ret["lerr"] = abs(float(S / N) - float(loss["rgb"])) / float(loss["rgb"])
            ret["shard"] = (nb, poff, rows, roff)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,case", [
    ("disk_rows", dict(H=24, W=32, patch_H=12, patch_W=16, batch_size=3, layers=(None, 32, 32, 3), L_2D=3, use_masks=True)),
    ("nomask_patches", dict(H=24, W=32, patch_H=12, patch_W=16, batch_size=4, layers=(None, 32, 32, 3), L_2D=3, use_masks=False)),
    ("implicit_patches", dict(H=24, W=32, patch_H=12, patch_W=16, batch_size=2, layers=(None, 32, 32, 3), L_2D=3,
                              use_masks=True, use_implicit_mask=True)),
    ("implicit_rows", dict(H=24, W=32, patch_H=12, patch_W=16, batch_size=3, layers=(None, 32, 32, 3), L_2D=3,
                           use_masks=True, use_implicit_mask=True)),
])
def test_two_rank_exchange_reproduces_single_rank(name, case):
    port = 29600 + (abs(hash(name)) % 300)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, case, ret), nprocs=2, join=True)
        assert ret["gerr"] < 1e-7, dict(ret)
        assert ret["lerr"] < 1e-9, dict(ret)
