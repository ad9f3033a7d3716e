"""Synthetic planar scenes (setup-time).

The upstream-BARF patch generator was removed from the reference (vestiges: `warp.noise_h/noise_t`,
options/planar.yaml:62-63; `self.warp_pert = None`, model/planar.py:55), so synthetic configs need their own.
The canvas is an analytic band-limited colour field, so a patch under a ground-truth sl(3) warp is evaluated
exactly at the warped pixel grid (no resampling), quantised to 8 bits like the reference's PNG inputs.
Everything here is data preparation outside the training step.
"""
import os

import numpy as np
import torch

from . import inputs
from .attrdict import AttrDict as edict
from .warp import Warp


def _field(xy, coef):
    """xy [...,2] (normalised canvas coords) -> rgb [...,3] in [0,1]."""
    fx, fy, ph, amp = coef            # each [3,K]
    x, y = xy[..., 0:1, None], xy[..., 1:2, None]          # [...,1,1]
    s = (amp * torch.sin(fx * x + fy * y + ph)).sum(-1)    # [...,3]
    return (0.5 + 0.25 * s).clamp(0, 1)


def make_coef(seed, n_waves=24, max_freq=40.0, device="cpu"):
    rs = np.random.RandomState(seed)
    fx = rs.uniform(-max_freq, max_freq, size=(3, n_waves))
    fy = rs.uniform(-max_freq, max_freq, size=(3, n_waves))
    ph = rs.uniform(0, 2 * np.pi, size=(3, n_waves))
    amp = rs.uniform(0.2, 1.0, size=(3, n_waves)) / np.sqrt(n_waves / 6.0)
    return tuple(torch.tensor(a, dtype=torch.float32, device=device) for a in (fx, fy, ph, amp))


def make_gt_warp(seed, B, noise_h, noise_t, device="cpu"):
    """Ground-truth sl(3) parameters: N(0, noise_h²) on the 6 non-translation dofs (scaled by 0.5 to keep patches
    mostly inside the canvas), uniform ±noise_t translations, patch 0 unperturbed (`warp.fix_first`)."""
    rs = np.random.RandomState(seed + 7919)
    h = rs.normal(0, 0.5 * noise_h, size=(B, 8))
    h[:, :2] = rs.uniform(-noise_t, noise_t, size=(B, 2))
    h[0] = 0
    return torch.tensor(h, dtype=torch.float32, device=device)


def make_scene(opt, seed=0, occluders=False, device=None):
    """-> images container with the same keys as inputs.prepare_images plus `gt_warp` [B,8]."""
    device = device or opt.device
    B, h, w = opt.batch_size, (opt.patch_H if opt.use_cropped_images else opt.H), (opt.patch_W if opt.use_cropped_images else opt.W)
    coef = make_coef(seed, device=device)
    wp = Warp(opt)
    gt_warp = make_gt_warp(seed, B, float(opt.warp.noise_h), float(opt.warp.noise_t), device=device)
    rs = np.random.RandomState(seed + 104729)
    rgbs, masks = [], []
    grid1 = wp.get_normalized_pixel_grid(crop=bool(opt.use_cropped_images))[:1].to(device)       # [1,P,2]
    for b in range(B):                      # patch at a time: large configs never hold [B,P,2,K] temporaries
        # a patch pixel shows the canvas at the INVERSE of the warp the optimiser must find: f(W_b^{-1}... the
        # reference convention is image_b(x) ~ f(W_b(x)) (model/planar.py:333-334), so sample at W_b(grid).
        xy = wp.warp_grid(grid1.contiguous(), gt_warp[b:b + 1].contiguous())
        chunks = [_field(c, coef) for c in xy[0].split(1 << 18)]
        patch = torch.cat(chunks).view(h, w, 3).permute(2, 0, 1)
        m = torch.ones(1, h, w, device=device)
        if occluders:
            for _ in range(rs.randint(1, 4)):
                y0, x0 = rs.randint(0, max(1, h - h // 4)), rs.randint(0, max(1, w - w // 4))
                hh, ww = rs.randint(h // 8, h // 3), rs.randint(w // 8, w // 3)
                colour = torch.tensor(rs.uniform(0, 1, size=3), dtype=torch.float32, device=device)
                patch[:, y0:y0 + hh, x0:x0 + ww] = colour[:, None, None]
                m[:, y0:y0 + hh, x0:x0 + ww] = 0
        rgbs.append(torch.round(patch * 255) / 255)
        masks.append(m)
    out = edict()
    out.rgb = torch.stack(rgbs).contiguous()
    out.gt_warp = gt_warp
    out.gt_hom = None
    full = Warp(opt).get_normalized_pixel_grid(crop=False)[0].to(device)
    out.gt = torch.cat([_field(c, coef) for c in full.split(1 << 18)]).view(opt.H, opt.W, 3).permute(2, 0, 1).contiguous()
    out.masks = torch.stack(masks).contiguous() if opt.use_masks else None
    out.masks_eroded = inputs.erode_images(out.masks, device, kernel=(5, 5)) if out.masks is not None else None
    gray = (0.299 * out.rgb[:, 0:1] + 0.587 * out.rgb[:, 1:2] + 0.114 * out.rgb[:, 2:3])
    out.gray = torch.round(gray * 255) / 255
    out.edges = inputs.compute_edges(out.gray, device) if opt.use_edges else None
    return out


def write_dataset(images, path):
    """Dump a scene in the reference's on-disk layout (i.png, i-m.png with white = occluded, gt.png) so the
    unmodified reference can read the same bytes (inputs.py:16-33, model/planar.py:62-70)."""
    import PIL.Image
    os.makedirs(path, exist_ok=True)

    def png(t):
        a = (t.detach().cpu().clamp(0, 1) * 255).round().byte().permute(1, 2, 0).numpy()
        return PIL.Image.fromarray(a[:, :, 0] if a.shape[2] == 1 else a)
    for i, im in enumerate(images.rgb):
        png(im).save(os.path.join(path, f"{i}.png"))
        if images.get("masks") is not None:
            png(1 - images.masks[i]).save(os.path.join(path, f"{i}-m.png"))
    png(images.gt).save(os.path.join(path, "gt.png"))
    np.savetxt(os.path.join(path, "gt_warp.txt"), images.gt_warp.cpu().numpy())
