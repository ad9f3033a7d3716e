"""CPU oracle for MARF's planar bundle-adjusting training step.

TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  The product path (`marf_b200`) never
routes through it and has no CPU fallback.

It is a functional restatement (torch, CPU, autograd for the gradients) of the
reference's eager-PyTorch hot path.  Every function cites the reference
file:line it follows (paths relative to the upstream repository root).

Parity status: PINNED.  `oracle/gen_golden.py` runs the *unmodified* reference
modules (imported from /root/reference with stub packages for its missing pip
dependencies) on seeded inputs and stores their outputs under `tests/golden/`;
`tests/test_oracle_vs_golden.py` checks every function here against those
vectors.  The reference itself ships no tests or golden vectors.

Third-party arithmetic the reference relies on (not present in its tree):
PyTorch (`requirements.yaml`: pytorch>=1.9; run here with torch 2.11.0) for
`matrix_exp`, `nn.Linear`, `sin/cos`, `sigmoid`; OpenCV (4.13 here) for the
Sobel / Gaussian / erode of the edge branch.  The oracle calls the same torch
ops on the CPU, and restates the OpenCV stencils in numpy (`sobel_gauss_edges`,
pinned against `cv2` in the golden tests).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# --------------------------------------------------------------------------- config
@dataclass
class PlanarConfig:
    """The subset of `options/planar.yaml` keys the hot path consumes."""
    H: int = 360
    W: int = 480
    patch_H: int = 180
    patch_W: int = 240
    batch_size: int = 5
    layers: Sequence[Optional[int]] = (None, 256, 256, 256, 256, 3)
    skip: Sequence[int] = ()
    L_2D: Optional[int] = 8            # None <=> `--arch.posenc!`
    barf_c2f: Optional[Tuple[float, float]] = None
    use_masks: bool = True
    use_implicit_mask: bool = False
    use_edges: bool = False
    alpha_initial: float = 0.0
    alpha_final: float = 1.0
    use_cropped_images: bool = True
    max_iter: int = 3000
    # loss weights in log10 scale, None = term skipped (options/planar.yaml:67-71)
    loss_weight: Dict[str, Optional[float]] = field(
        default_factory=lambda: dict(render=0, rgb=0, edge=0, mask=0))

    @property
    def d_in(self) -> int:
        return 2 + 4 * self.L_2D if self.L_2D else 2

    @property
    def h(self) -> int:
        return self.patch_H if self.use_cropped_images else self.H

    @property
    def w(self) -> int:
        return self.patch_W if self.use_cropped_images else self.W


# --------------------------------------------------------------------------- geometry
def crop_window(cfg: PlanarConfig) -> Tuple[Tuple[int, int], Tuple[int, int], float, float]:
    """warp.py:9-21 — crop ranges and the aspect normalisers."""
    y_crop = (cfg.H // 2 - cfg.patch_H // 2, cfg.H // 2 + cfg.patch_H // 2)
    x_crop = (cfg.W // 2 - cfg.patch_W // 2, cfg.W // 2 + cfg.patch_W // 2)
    norm_h = cfg.H / max(cfg.H, cfg.W)
    norm_w = cfg.W / max(cfg.H, cfg.W)
    return y_crop, x_crop, norm_h, norm_w


def normalized_pixel_grid(cfg: PlanarConfig, crop: bool = True, dtype=torch.float32) -> torch.Tensor:
    """warp.py:33-68 — [P,2] grid in (x,y) order, row-major over (row,col).

    The op order ((i+0.5)/n*2-1)*norm is kept so the f32 roundings agree."""
    y_crop, x_crop, norm_h, norm_w = crop_window(cfg)
    if crop:
        ys = torch.arange(*y_crop, dtype=dtype)
        xs = torch.arange(*x_crop, dtype=dtype)
    else:
        ys = torch.arange(cfg.H, dtype=dtype)
        xs = torch.arange(cfg.W, dtype=dtype)
    y_range = ((ys + 0.5) / cfg.H * 2 - 1) * norm_h
    x_range = ((xs + 0.5) / cfg.W * 2 - 1) * norm_w
    Y, X = torch.meshgrid(y_range, x_range, indexing="ij")
    return torch.stack([X, Y], dim=-1).reshape(-1, 2)


def sl3_to_SL3(h: torch.Tensor) -> torch.Tensor:
    """warp.py:98-106 — 8-vector -> traceless 3x3 generator -> matrix exponential."""
    h1, h2, h3, h4, h5, h6, h7, h8 = [h[..., i] for i in range(8)]
    A = torch.stack([
        torch.stack([h5, h3, h1], dim=-1),
        torch.stack([h4, -h5 - h6, h2], dim=-1),
        torch.stack([h7, h8, h6], dim=-1)], dim=-2)
    return torch.linalg.matrix_exp(A)


def warp_grid(xy: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """warp.py:27-31,70-81 — q=[x,y,1]·Hᵀ ; (u,v)=q_xy/(q_z+1e-8).  xy [B,P,2], h [B,8]."""
    hom = torch.cat([xy, torch.ones_like(xy[..., :1])], dim=-1)
    q = hom @ sl3_to_SL3(h).transpose(-2, -1)
    return q[..., :2] / (q[..., 2:] + 1e-8)


def warp_corners(cfg: PlanarConfig, h: torch.Tensor) -> torch.Tensor:
    """warp.py:83-93 — the four crop corners pushed through each patch's warp, [B,4,2]."""
    y_crop, x_crop, norm_h, norm_w = crop_window(cfg)
    Y = [((y + 0.5) / cfg.H * 2 - 1) * norm_h for y in y_crop]
    X = [((x + 0.5) / cfg.W * 2 - 1) * norm_w for x in x_crop]
    corners = torch.tensor([(X[0], Y[0]), (X[0], Y[1]), (X[1], Y[1]), (X[1], Y[0])],
                           dtype=h.dtype).repeat(h.shape[0], 1, 1)
    return warp_grid(corners, h)


# --------------------------------------------------------------------------- neural image
def c2f_weights(L: int, barf_c2f, progress: float, dtype=torch.float32) -> torch.Tensor:
    """model/planar.py:462-467 — w_k = (1-cos(clamp(a-k,0,1)·π))/2, a=(progress-start)/(end-start)·L."""
    if barf_c2f is None:
        return torch.ones(L, dtype=dtype)
    start, end = barf_c2f
    alpha = (torch.tensor(progress, dtype=dtype) - start) / (end - start) * L
    k = torch.arange(L, dtype=dtype)
    return (1 - (alpha - k).clamp(min=0, max=1).mul(np.pi).cos()) / 2


def positional_encoding(coord: torch.Tensor, L: int, barf_c2f=None, progress: float = 0.0) -> torch.Tensor:
    """model/planar.py:451-471 — [..,2] -> [..,4L] ordered [sin u(L), cos u(L), sin v(L), cos v(L)]."""
    shape = coord.shape
    freq = 2 ** torch.arange(L, dtype=torch.float32) * np.pi
    freq = freq.to(coord.dtype)
    spectrum = coord[..., None] * freq
    enc = torch.stack([spectrum.sin(), spectrum.cos()], dim=-2).reshape(*shape[:-1], -1)
    if barf_c2f is not None:
        w = c2f_weights(L, barf_c2f, progress, dtype=coord.dtype)
        enc = (enc.reshape(-1, L) * w).reshape(*shape[:-1], -1)
    return enc


def neural_image(coord: torch.Tensor, weights: List[torch.Tensor], biases: List[torch.Tensor],
                 cfg: PlanarConfig, progress: float = 0.0) -> torch.Tensor:
    """model/planar.py:429-449 — posenc ⊕ xy -> ReLU MLP (optional skip concat) -> sigmoid."""
    if cfg.L_2D:
        enc = positional_encoding(coord, cfg.L_2D, cfg.barf_c2f, progress)
        points = torch.cat([coord, enc], dim=-1)
    else:
        points = coord
    feat = points
    n = len(weights)
    for li in range(n):
        if li in cfg.skip:
            feat = torch.cat([feat, points], dim=-1)
        feat = torch.nn.functional.linear(feat, weights[li], biases[li])
        if li != n - 1:
            feat = torch.relu(feat)
    return torch.sigmoid(feat)


def layer_shapes(cfg: PlanarConfig) -> List[Tuple[int, int]]:
    """model/planar.py:410-421 + util.py:105-108 — (k_out,k_in) per Linear."""
    dims = list(zip(cfg.layers[:-1], cfg.layers[1:]))
    out = []
    for li, (k_in, k_out) in enumerate(dims):
        if li == 0:
            k_in = cfg.d_in
        if li in cfg.skip:
            k_in += cfg.d_in
        out.append((k_out, k_in))
    return out


# --------------------------------------------------------------------------- implicit mask
MASK_UV_FREQS = 10          # PosEmbedding(10-1, 10): model/planar.py:320
MASK_EMBED_DIM = 128        # nn.Embedding(N_vocab,128): model/planar.py:327
MASK_IN = 3 * MASK_EMBED_DIM + 2 + 2 * 2 * MASK_UV_FREQS   # 426


def pos_embedding(x: torch.Tensor, n_freqs: int = MASK_UV_FREQS) -> torch.Tensor:
    """model/planar.py:491-518 — [x, sin(2^k x), cos(2^k x) ...], k=0..n-1, no π, per-frequency interleave."""
    freqs = 2 ** torch.linspace(0, n_freqs - 1, n_freqs)
    out = [x]
    for f in freqs:
        out += [torch.sin(f.to(x.dtype) * x), torch.cos(f.to(x.dtype) * x)]
    return torch.cat(out, dim=-1)


def mask_features(image: torch.Tensor, xy0: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """model/planar.py:342-349 — per patch: trunc(rgb) indices -> colour embedding (R,G,B blocks) ⊕ uv embedding.

    image [3,h,w] f32, xy0 [P,2] (UN-warped grid), embed [N_vocab,128] -> [P,426]."""
    idx = image.long().reshape(3, -1).permute(1, 0)                 # [P,3] ∈ {0,1}
    col = embed[idx].reshape(idx.shape[0], 3 * embed.shape[1])      # [P,384]
    return torch.cat([col, pos_embedding(xy0)], dim=-1)


def mask_head(feats: torch.Tensor, weights: List[torch.Tensor], biases: List[torch.Tensor]) -> torch.Tensor:
    """model/planar.py:475-488 — 426→256→256→256→256→1, ReLU ×4, sigmoid."""
    x = feats
    n = len(weights)
    for li in range(n):
        x = torch.nn.functional.linear(x, weights[li], biases[li])
        x = torch.relu(x) if li != n - 1 else torch.sigmoid(x)
    return x


# --------------------------------------------------------------------------- edge branch
def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    """OpenCV BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba)."""
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def _sep_filter(img: np.ndarray, ky: np.ndarray, kx: np.ndarray) -> np.ndarray:
    """Separable correlation in float64 with BORDER_REFLECT_101.  img [h,w,c]."""
    h, w = img.shape[:2]
    ry, rx = len(ky) // 2, len(kx) // 2
    tmp = np.zeros_like(img, dtype=np.float64)
    cols = np.arange(w)
    for t, k in enumerate(kx):
        tmp += k * img[:, _reflect101(cols + t - rx, w)]
    out = np.zeros_like(tmp)
    rows = np.arange(h)
    for t, k in enumerate(ky):
        out += k * tmp[_reflect101(rows + t - ry, h)]
    return out


GAUSS5 = np.array([1.0, 4.0, 6.0, 4.0, 1.0]) / 16.0   # cv2.getGaussianKernel(5, 0): fixed table for ksize<=7, sigma<=0


def sobel_gauss_edges(images: np.ndarray) -> np.ndarray:
    """inputs.py:50-69 restated without OpenCV: per image, Sobel-3 in x and y (float64,
    default border = REFLECT_101), magnitude, then 5x5 Gaussian blur with sigma=0.

    images [B,C,h,w] (any float dtype) -> [B,C,h,w] float64."""
    out = []
    d = np.array([-1.0, 0.0, 1.0])
    s = np.array([1.0, 2.0, 1.0])
    for im in images:
        x = np.transpose(np.asarray(im), (1, 2, 0)).astype(np.float64)
        gx = _sep_filter(x, ky=s, kx=d)
        gy = _sep_filter(x, ky=d, kx=s)
        mag = np.sqrt(gx ** 2 + gy ** 2)
        blur = _sep_filter(mag, ky=GAUSS5, kx=GAUSS5)
        out.append(np.transpose(blur, (2, 0, 1)))
    return np.stack(out)


def compute_edges_cv2(images: torch.Tensor) -> torch.Tensor:
    """inputs.py:50-69 with the same OpenCV calls the reference makes (detached, float64)."""
    import cv2
    res = []
    for image in images:
        i = np.transpose(image.detach().cpu().numpy(), (1, 2, 0))
        sx = cv2.Sobel(i, cv2.CV_64F, 1, 0, ksize=3)
        sy = cv2.Sobel(i, cv2.CV_64F, 0, 1, ksize=3)
        i = np.sqrt(sx ** 2 + sy ** 2)
        i = cv2.GaussianBlur(i, (5, 5), 0)
        if i.ndim == 2:
            i = i[:, :, None]
        res.append(torch.from_numpy(np.ascontiguousarray(np.transpose(i, (2, 0, 1)))))
    return torch.stack(res)


def erode5(masks: np.ndarray) -> np.ndarray:
    """inputs.py:71-85 — 5x5 rectangular erosion (cv2.erode default border = +inf constant).  [B,1,h,w]."""
    B, C, h, w = masks.shape
    pad = np.pad(masks, ((0, 0), (0, 0), (2, 2), (2, 2)), constant_values=np.inf)
    out = np.full_like(masks, np.inf)
    for dy in range(5):
        for dx in range(5):
            out = np.minimum(out, pad[:, :, dy:dy + h, dx:dx + w])
    return out


# --------------------------------------------------------------------------- loss
def mse_loss(pred: torch.Tensor, labels: torch.Tensor, masks: Optional[torch.Tensor] = None) -> torch.Tensor:
    """model/planar.py:382-391 — mean((p-l)²) or Σ((p-l)·m)²/(3·Σm) (mask squared in the numerator only)."""
    if masks is None:
        return ((pred - labels) ** 2).mean()
    return (((pred - labels) * masks) ** 2).sum() / (masks.sum() * 3)


def edge_alpha(cfg: PlanarConfig, it: int) -> float:
    """model/planar.py:359."""
    if not cfg.use_edges:
        return 0
    return cfg.alpha_initial + (cfg.alpha_final - cfg.alpha_initial) * (it / cfg.max_iter)


def loss_coefficients(cfg: PlanarConfig, it: int) -> Tuple[float, float, float]:
    """Coefficients of (rgb, mask, edge) inside `loss.all` once `render` is expanded:
    model/planar.py:371-378 (render = (1-α)rgb + 0.5 mask + α edge) and :177-184 (Σ 10^w · loss)."""
    a = edge_alpha(cfg, it)
    lw = cfg.loss_weight

    def p(key):
        return 0.0 if lw.get(key) is None else 10 ** float(lw[key])
    if lw.get("render", 0) is None:
        # the reference builds no loss entries at all when render is None (model/planar.py:361)
        return 0.0, 0.0, 0.0
    return p("render") * (1 - a) + p("rgb"), p("render") * 0.5 + p("mask"), p("render") * a + p("edge")


# --------------------------------------------------------------------------- the step
@dataclass
class PlanarParams:
    mlp_w: List[torch.Tensor]
    mlp_b: List[torch.Tensor]
    warp: torch.Tensor                              # [B,8]
    mask_w: Optional[List[torch.Tensor]] = None
    mask_b: Optional[List[torch.Tensor]] = None
    embed: Optional[torch.Tensor] = None            # embedding_view.weight [N_vocab,128]

    def leaves(self) -> List[torch.Tensor]:
        out = list(self.mlp_w) + list(self.mlp_b) + [self.warp]
        if self.mask_w is not None:
            out += list(self.mask_w) + list(self.mask_b)
        return out


def forward(params: PlanarParams, images_rgb: torch.Tensor, cfg: PlanarConfig, progress: float):
    """model/planar.py:329-353 (Graph.forward).  Returns dict with rgb_prediction [B,P,3],
    rgb_prediction_map [B,3,h,w], edge_prediction (f64, detached) and, with the implicit
    mask on, mask_prediction [B,P,1] / mask_prediction_map [B,1,h,w]."""
    B, h, w = cfg.batch_size, cfg.h, cfg.w
    xy = normalized_pixel_grid(cfg, crop=cfg.use_cropped_images).to(params.warp.dtype)
    xy_b = xy.repeat(B, 1, 1)
    uv = warp_grid(xy_b, params.warp)
    rgb = neural_image(uv, params.mlp_w, params.mlp_b, cfg, progress)
    out = dict(rgb_prediction=rgb, rgb_prediction_map=rgb.view(B, h, w, 3).permute(0, 3, 1, 2))
    out["edge_prediction"] = compute_edges_cv2(out["rgb_prediction_map"]).to(rgb.device)   # (the reference: .to(opt.device))
    if cfg.use_implicit_mask:
        preds = []
        for im in images_rgb:
            feats = mask_features(im, xy, params.embed)
            preds.append(mask_head(feats, params.mask_w, params.mask_b))
        m = torch.stack(preds)
        out["mask_prediction"] = m
        out["mask_prediction_map"] = m.view(B, h, w, 1).permute(0, 3, 1, 2)
    return out


def losses(out: dict, images: dict, cfg: PlanarConfig, it: int) -> Dict[str, torch.Tensor]:
    """model/planar.py:355-380 (Graph.compute_loss) + :172-185 (Model.summarize_loss)."""
    loss: Dict[str, torch.Tensor] = {}
    alpha = edge_alpha(cfg, it)
    if cfg.loss_weight.get("render", 0) is not None:
        if cfg.use_implicit_mask:
            m_rgb = m_edge = out["mask_prediction_map"]
        else:
            m_rgb, m_edge = images.get("masks"), images.get("masks_eroded")
        rgb_loss = mse_loss(out["rgb_prediction_map"], images["rgb"], m_rgb)
        edge_loss = mse_loss(out["edge_prediction"], images["edges"], m_edge) if cfg.use_edges else torch.tensor(0)
        mask_loss = ((1 - out["mask_prediction_map"]) ** 2).mean() if cfg.use_implicit_mask else torch.tensor(0)
        loss["render"] = (1 - alpha) * rgb_loss + 0.5 * mask_loss + alpha * edge_loss
        loss["rgb"], loss["mask"], loss["edge"] = rgb_loss, mask_loss, edge_loss
    total = 0.0
    for key, val in loss.items():
        if cfg.loss_weight.get(key) is not None:
            total = total + 10 ** float(cfg.loss_weight[key]) * val
    loss["all"] = total
    return loss


def step(params: PlanarParams, images: dict, cfg: PlanarConfig, it: int = 0, progress: Optional[float] = None):
    """One forward + loss + backward of the reference's train_iteration (model/planar.py:192-196),
    without the optimizer.  Returns (outputs, losses, grads) — grads in `params.leaves()` order."""
    if progress is None:
        progress = it / cfg.max_iter
    leaves = params.leaves()
    for t in leaves:
        t.requires_grad_(True)
        t.grad = None
    out = forward(params, images["rgb"], cfg, progress)
    loss = losses(out, images, cfg, it)
    loss["all"].backward()
    grads = [t.grad if t.grad is not None else torch.zeros_like(t) for t in leaves]
    return out, loss, grads


# --------------------------------------------------------------------------- initialisation
def init_params(cfg: PlanarConfig, seed: int, n_vocab: int = 1500, dtype=torch.float32) -> PlanarParams:
    """Seed-matched construction in the reference's RNG order: the image MLP's
    nn.Linear layers (model/planar.py:414-427, with the ×sqrt(D_in/2) first-layer rescale under c2f),
    warp embedding N(0,1)→zeroed (:310-311), mask head Linear layers (:326, :480-484), embedding_view (:327)."""
    torch.manual_seed(seed)
    mlp_w, mlp_b = [], []
    for li, (k_out, k_in) in enumerate(layer_shapes(cfg)):
        lin = torch.nn.Linear(k_in, k_out)
        if cfg.barf_c2f is not None and li == 0:
            scale = np.sqrt(cfg.d_in / 2.)
            lin.weight.data *= scale
            lin.bias.data *= scale
        mlp_w.append(lin.weight.data.to(dtype))
        mlp_b.append(lin.bias.data.to(dtype))
    warp = torch.nn.Embedding(cfg.batch_size, 8).weight.data
    warp = torch.zeros_like(warp).to(dtype)
    p = PlanarParams(mlp_w, mlp_b, warp)
    if cfg.use_implicit_mask:
        dims = [(256, MASK_IN), (256, 256), (256, 256), (256, 256), (1, 256)]
        p.mask_w, p.mask_b = [], []
        for k_out, k_in in dims:
            lin = torch.nn.Linear(k_in, k_out)
            p.mask_w.append(lin.weight.data.to(dtype))
            p.mask_b.append(lin.bias.data.to(dtype))
        p.embed = torch.nn.Embedding(n_vocab, MASK_EMBED_DIM).weight.data.to(dtype)
    return p


def adam_train(params: PlanarParams, images: dict, cfg: PlanarConfig, n_iter: int,
               lr=1e-3, lr_warp=1e-3, lr_mask=1e-3, fix_first=True):
    """The reference's training loop restricted to what changes state (model/planar.py:154-158,187-209):
    Adam over {mlp}, {warp}, [{mask head}]; `fix_first` zeroes warp row 0 after every step;
    progress <- it/max_iter after every step.  Returns per-iteration loss dicts (floats)."""
    groups = [dict(params=list(params.mlp_w) + list(params.mlp_b), lr=lr), dict(params=[params.warp], lr=lr_warp)]
    if cfg.use_implicit_mask:
        groups.append(dict(params=list(params.mask_w) + list(params.mask_b), lr=lr_mask))
    for g in groups:
        for t in g["params"]:
            t.requires_grad_(True)
    opt = torch.optim.Adam(groups)
    history = []
    progress = 0.0
    for it in range(n_iter):
        opt.zero_grad()
        out = forward(params, images["rgb"], cfg, progress)
        loss = losses(out, images, cfg, it)
        loss["all"].backward()
        opt.step()
        if fix_first:
            params.warp.data[0] = 0
        progress = (it + 1) / cfg.max_iter
        history.append({k: float(v) for k, v in loss.items()})
    return history
