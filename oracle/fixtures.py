"""Deterministic synthetic inputs shared by `oracle/gen_golden.py` and the tests.

TEST INFRASTRUCTURE ONLY.  Everything is numpy `RandomState` + 8-bit quantisation, so the
same bytes come out in the build container (where the goldens are made from the reference)
and on the GPU box (where the CUDA path is checked against them).
"""
import numpy as np
import torch


def synth_patches(seed: int, B: int, h: int, w: int, occluders: bool = True):
    """Band-limited colour patches in [0,1] quantised to k/255 (like the reference's PNG inputs),
    with a few saturated (==1.0) blobs so `trunc(rgb)` hits both embedding rows, plus binary
    validity masks (1 = valid, the convention after inputs.py:30-31)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, h), np.linspace(-1, 1, w), indexing="ij")
    rgb = np.zeros((B, 3, h, w), dtype=np.float64)
    for b in range(B):
        for c in range(3):
            acc = np.zeros((h, w))
            for _ in range(6):
                fx, fy = rs.uniform(-6, 6, size=2)
                ph = rs.uniform(0, 2 * np.pi)
                acc += rs.uniform(0.2, 1.0) * np.sin(fx * xx + fy * yy + ph)
            rgb[b, c] = 0.5 + 0.22 * acc
    rgb = np.clip(rgb, 0, 1)
    masks = np.ones((B, 1, h, w), dtype=np.float32)
    for b in range(B):
        # a saturated blob (all channels 1.0)
        cy, cx, r = rs.uniform(-0.6, 0.6), rs.uniform(-0.6, 0.6), rs.uniform(0.08, 0.2)
        blob = (yy - cy) ** 2 + (xx - cx) ** 2 < r ** 2
        rgb[b][:, blob] = 1.0
        # single-channel saturation
        cy, cx, r = rs.uniform(-0.6, 0.6), rs.uniform(-0.6, 0.6), rs.uniform(0.08, 0.2)
        blob = (yy - cy) ** 2 + (xx - cx) ** 2 < r ** 2
        rgb[b, b % 3][blob] = 1.0
        if occluders:
            y0, x0 = rs.uniform(-0.8, 0.3), rs.uniform(-0.8, 0.3)
            box = (yy > y0) & (yy < y0 + rs.uniform(0.2, 0.5)) & (xx > x0) & (xx < x0 + rs.uniform(0.2, 0.5))
            masks[b, 0][box] = 0.0
            rgb[b][:, box] = rs.uniform(0, 1, size=3)[:, None]
    rgb = np.round(rgb * 255).astype(np.uint8).astype(np.float32) / 255.0
    return torch.from_numpy(rgb), torch.from_numpy(masks)


def synth_mlp(seed: int, shapes, scale: float = 1.0):
    """nn.Linear-like uniform(-1/sqrt(k_in), 1/sqrt(k_in)) weights from numpy's RandomState.
    shapes: list of (k_out,k_in)."""
    rs = np.random.RandomState(seed)
    ws, bs = [], []
    for k_out, k_in in shapes:
        bound = 1.0 / np.sqrt(k_in)
        ws.append(torch.from_numpy(rs.uniform(-bound, bound, size=(k_out, k_in)).astype(np.float32) * scale))
        bs.append(torch.from_numpy(rs.uniform(-bound, bound, size=(k_out,)).astype(np.float32) * scale))
    return ws, bs


def synth_warp(seed: int, B: int, scale: float = 0.05, fix_first: bool = True):
    rs = np.random.RandomState(seed)
    h = rs.normal(0, scale, size=(B, 8)).astype(np.float32)
    if fix_first:
        h[0] = 0
    return torch.from_numpy(h)


def synth_embed(seed: int, n_vocab: int = 1500, dim: int = 128):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.normal(0, 1, size=(n_vocab, dim)).astype(np.float32))


def digest(t, n_probe: int = 64, seed: int = 1234):
    """Compact pin for a big tensor: norm, sum, projection on a seeded ±1 vector, a strided sample."""
    a = np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float64).ravel()
    rs = np.random.RandomState(seed)
    sign = rs.randint(0, 2, size=a.size) * 2.0 - 1.0
    idx = np.linspace(0, a.size - 1, min(n_probe, a.size)).astype(np.int64)
    return dict(norm=float(np.linalg.norm(a)), sum=float(a.sum()), proj=float((a * sign).sum()),
                sample=a[idx].astype(np.float64), size=int(a.size))
