"""Helpers for the GPU parity tests: run a named case through the C ABI (marf_b200.engine.PlanarEngine)."""
import torch

import planar_oracle as po
from marf_b200 import _lib as L
from marf_b200.engine import PlanarEngine


def mask_mode_of(cfg: po.PlanarConfig):
    if cfg.use_implicit_mask:
        return L.MASK_IMPLICIT
    return L.MASK_DISK if cfg.use_masks else L.MASK_NONE


def make_engine(cfg: po.PlanarConfig, precision="fp32", rank=0, world=1, max_chunk_pixels=0):
    return PlanarEngine(H=cfg.H, W=cfg.W, patch_H=cfg.patch_H, patch_W=cfg.patch_W, batch_size=cfg.batch_size,
                        layers=list(cfg.layers[1:]), skip=list(cfg.skip), L_2D=cfg.L_2D, barf_c2f=cfg.barf_c2f,
                        mask_mode=mask_mode_of(cfg), use_edges=cfg.use_edges, use_cropped=cfg.use_cropped_images,
                        precision=precision, device="cuda:0", rank=rank, world=world, max_chunk_pixels=max_chunk_pixels)


def step_kwargs(eng: PlanarEngine, cfg: po.PlanarConfig, params: po.PlanarParams, images: dict, it: int, progress: float):
    """Device copies of a case's parameters / inputs and fresh output tensors: the keyword arguments of PlanarEngine.step."""
    dev = eng.device
    cu = lambda t: None if t is None else t.detach().to(dev).contiguous()
    mlp_w = [cu(t) for t in params.mlp_w]
    mlp_b = [cu(t) for t in params.mlp_b]
    warp = cu(params.warp)
    implicit = cfg.use_implicit_mask
    mask_w = [cu(t) for t in params.mask_w] if implicit else None
    mask_b = [cu(t) for t in params.mask_b] if implicit else None
    embed = cu(params.embed) if implicit else None
    B, h, w = cfg.batch_size, cfg.h, cfg.w
    g_mlp_w = [torch.full_like(t, 7.0) for t in mlp_w]
    g_mlp_b = [torch.full_like(t, 7.0) for t in mlp_b]
    g_warp = torch.full_like(warp, 7.0)
    g_mask_w = [torch.full_like(t, 7.0) for t in mask_w] if implicit else None
    g_mask_b = [torch.full_like(t, 7.0) for t in mask_b] if implicit else None
    rgb_pred = torch.zeros(B, h * w, 3, device=dev)
    mask_pred = torch.zeros(B, h * w, 1, device=dev) if implicit else None
    edge_pred = torch.zeros(B, 3, h, w, dtype=torch.float64, device=dev) if cfg.use_edges else None
    coef = po.loss_coefficients(cfg, it)
    return dict(mlp_w=mlp_w, mlp_b=mlp_b, warp=warp, rgb=cu(images["rgb"]),
                masks=cu(images["masks"]) if cfg.use_masks and not implicit else None,
                masks_eroded=cu(images["masks_eroded"]) if cfg.use_masks and not implicit and cfg.use_edges else None,
                edges=cu(images["edges"]) if cfg.use_edges else None,
                mask_w=mask_w, mask_b=mask_b, embed=embed, g_mlp_w=g_mlp_w, g_mlp_b=g_mlp_b, g_warp=g_warp,
                g_mask_w=g_mask_w, g_mask_b=g_mask_b, rgb_pred=rgb_pred, mask_pred=mask_pred, edge_pred=edge_pred,
                progress=progress, coef=coef)


def run_step(eng: PlanarEngine, cfg: po.PlanarConfig, params: po.PlanarParams, images: dict, it: int, progress: float,
             two_phase=False, sync=True):
    """Returns dict(rgb_pred, mask_pred, edge_pred, grads{...}, losses{...}) on the CPU."""
    kw = step_kwargs(eng, cfg, params, images, it, progress)
    implicit = cfg.use_implicit_mask
    coef = kw["coef"]
    g_mlp_w, g_mlp_b, g_warp, g_mask_w, g_mask_b = kw["g_mlp_w"], kw["g_mlp_b"], kw["g_warp"], kw["g_mask_w"], kw["g_mask_b"]
    rgb_pred, mask_pred, edge_pred = kw["rgb_pred"], kw["mask_pred"], kw["edge_pred"]
    if two_phase:
        eng.step_forward(**kw)
        eng.step_backward()
    else:
        eng.step(**kw)
    torch.cuda.synchronize()
    rgb_l, mask_l, edge_l = [float(x) for x in eng.loss_values()]
    a = po.edge_alpha(cfg, it)
    losses = dict(rgb=rgb_l, mask=mask_l, edge=edge_l, render=(1 - a) * rgb_l + 0.5 * mask_l + a * edge_l,
                  all=coef[0] * rgb_l + coef[1] * mask_l + coef[2] * edge_l)
    grads = {}
    for i, (gw, gb) in enumerate(zip(g_mlp_w, g_mlp_b)):
        grads[f"gW{i}"], grads[f"gb{i}"] = gw.cpu(), gb.cpu()
    grads["gwarp"] = g_warp.cpu()
    if implicit:
        for i, (gw, gb) in enumerate(zip(g_mask_w, g_mask_b)):
            grads[f"gMW{i}"], grads[f"gMb{i}"] = gw.cpu(), gb.cpu()
    return dict(rgb_pred=rgb_pred.cpu(), mask_pred=None if mask_pred is None else mask_pred.cpu(),
                edge_pred=None if edge_pred is None else edge_pred.cpu(), grads=grads, losses=losses,
                nonfinite=float(eng.sums[L.NONFINITE]))
