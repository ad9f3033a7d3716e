"""Named test cases: the same seeded (config, parameters, images) that oracle/gen_golden.py fed
to the reference.  Shared by the oracle-vs-golden tests (CPU) and the CUDA parity tests (GPU)."""
import os

import numpy as np
import torch

import fixtures as fx
import planar_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SMALL = dict(H=40, W=56, patch_H=20, patch_W=28, batch_size=3, max_iter=200, layers=(None, 64, 64, 64, 3), L_2D=4)
MID = dict(H=72, W=96, patch_H=36, patch_W=48, batch_size=3, max_iter=3000)
FULL2 = dict(batch_size=2, max_iter=3000)

STEP_CASES = {
    "small_nomask": dict(SMALL, use_masks=False),
    "small_mask": dict(SMALL, use_masks=True),
    "small_mask_c2f": dict(SMALL, use_masks=True, barf_c2f=(0.0, 0.4)),
    "small_mask_edges": dict(SMALL, use_masks=True, use_edges=True),
    "small_noposenc": dict(SMALL, use_masks=True, L_2D=None),
    "small_skip": dict(SMALL, use_masks=True, skip=(2,)),
    "small_weights": dict(SMALL, use_masks=True, use_edges=True,
                          loss_weight=dict(render=0, rgb=-1, edge=0.5, mask=None)),
    "mid_mask": dict(MID, use_masks=True),
    "mid_mask_c2f": dict(MID, use_masks=True, barf_c2f=(0.0, 0.4)),
    "mid_nomask_edges": dict(MID, use_masks=False, use_edges=True),
    "implicit": dict(FULL2, use_masks=True, use_implicit_mask=True),
    "implicit_edges": dict(FULL2, use_masks=True, use_implicit_mask=True, use_edges=True),
}


# cases without a golden file (the oracle itself is pinned by the goldens of STEP_CASES): BASELINE config 5's network
ORACLE_ONLY_CASES = {
    "wide512_L10": dict(MID, use_masks=True, layers=(None, 512, 512, 512, 512, 3), L_2D=10),
    "wide512_c2f": dict(MID, use_masks=True, layers=(None, 512, 512, 512, 512, 3), L_2D=10, barf_c2f=(0.0, 0.4)),
    # BASELINE config 2's real shape: 5 patches of 180x240, learned mask + edge term, full posenc
    "implicit_edges_b5": dict(batch_size=5, max_iter=3000, use_masks=True, use_implicit_mask=True, use_edges=True),
}


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def make_images(cfg: po.PlanarConfig, seed: int):
    """Mirror of gen_golden.make_images, with the oracle's numpy stencils instead of OpenCV."""
    rgb, masks = fx.synth_patches(seed, cfg.batch_size, cfg.patch_H, cfg.patch_W, occluders=True)
    gray = (0.299 * rgb[:, 0:1] + 0.587 * rgb[:, 1:2] + 0.114 * rgb[:, 2:3])
    gray = torch.round(gray * 255) / 255
    im = dict(rgb=rgb, gray=gray, masks=None, masks_eroded=None, edges=None)
    if cfg.use_masks:
        im["masks"] = masks
        im["masks_eroded"] = torch.from_numpy(po.erode5(masks.numpy()))
    if cfg.use_edges:
        im["edges"] = torch.from_numpy(po.sobel_gauss_edges(gray.numpy()))
    return im


def build_case(name):
    """-> (cfg, params, images, it, progress, golden) for a step case."""
    cfg = po.PlanarConfig(**(STEP_CASES[name] if name in STEP_CASES else ORACLE_ONLY_CASES[name]))
    implicit = cfg.use_implicit_mask
    seed_w, seed_h, seed_im = (22, 32, 42) if implicit else (21, 31, 41)
    ws, bs = fx.synth_mlp(seed_w, po.layer_shapes(cfg), scale=2.0)
    params = po.PlanarParams(ws, bs, fx.synth_warp(seed_h, cfg.batch_size, scale=0.05))
    if implicit:
        mshapes = [(256, po.MASK_IN), (256, 256), (256, 256), (256, 256), (1, 256)]
        params.mask_w, params.mask_b = fx.synth_mlp(seed_w + 1, mshapes)
        params.embed = fx.synth_embed(seed_w + 2, 1500, 128)
    images = make_images(cfg, seed_im)
    if not os.path.exists(os.path.join(GOLDEN, "step_" + name + ".npz")):
        it = 450                                               # oracle-only case (the iteration gen_golden.py uses for MID)
        return cfg, params, images, it, it / cfg.max_iter, None
    g = load_golden("step_" + name)
    return cfg, params, images, int(g["it"]), float(g["progress"]), g


def check_digest(t, g, key, tol=1e-5):
    """Compare a tensor with a stored digest (norm / sum / ±1 projection / strided sample).
    `tol` is relative to the tensor's scale (max-abs of the sample for entries, the norm for sums)."""
    d = fx.digest(t)
    assert d["size"] == int(g[key + ".size"])
    norm = float(g[key + ".norm"])
    peak = float(np.abs(g[key + ".sample"]).max()) + 1e-12
    assert abs(d["norm"] - norm) <= tol * norm + 1e-12, (key, d["norm"], norm)
    # sums of n entries each off by ~tol*peak (random sign) -> sqrt(n) growth
    slack = 4 * tol * peak * np.sqrt(d["size"]) + 1e-12
    assert abs(d["sum"] - float(g[key + ".sum"])) <= slack, (key, d["sum"], float(g[key + ".sum"]), slack)
    assert abs(d["proj"] - float(g[key + ".proj"])) <= slack, (key, d["proj"], float(g[key + ".proj"]), slack)
    err = np.abs(d["sample"] - g[key + ".sample"]).max()
    assert err <= tol * peak, (key, err, tol * peak)


def check_close(a, b, tol, what=""):
    """max-abs error relative to the reference tensor's max-abs."""
    a = np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu().numpy() if hasattr(b, "detach") else b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    peak = np.abs(b).max() + 1e-12
    err = np.abs(a - b).max()
    assert err <= tol * peak, (what, err, tol * peak)
