"""Setup-time data path of the planar plugin (the reference's `inputs.py`): PNG patches, occlusion masks, eroded
masks, grey images, edge labels and ground-truth homographies -> the `images` container the step reads.

Same function names and return contract as the reference (inputs.py:16-127).  This runs once per training run,
outside the step; stencils run on the GPU through the C ABI (`marf_compute_edges`) or as
torch pooling, not through OpenCV.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .attrdict import AttrDict as edict


def _engine_for(device):
    from .warp import _util_engine
    return _util_engine(torch.device(device))


def _to_tensor(im):
    """PIL image -> float32 [C,h,w] in [0,1] (what torchvision's to_tensor does for 8-bit images)."""
    a = np.array(im)
    if a.ndim == 2:
        a = a[:, :, None]
    return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1))).float().div(255)


def load_images(fps, opt, mode="RGB", invert_gray=False):
    """inputs.py:16-33 — list of files -> [B,C,h,w]; LANCZOS thumbnail to the patch size; masks inverted so 1=valid."""
    import PIL.Image
    if not fps:
        return None
    if not isinstance(fps, list):
        raise TypeError("Function requires list of input filepaths!")
    out = []
    for fp in fps:
        im = PIL.Image.open(fp).convert(mode)
        if opt.use_cropped_images:
            im.thumbnail((opt.patch_W, opt.patch_H), PIL.Image.Resampling.LANCZOS)
        t = _to_tensor(im).to(opt.device)
        if mode == "L" and invert_gray:
            t = (t < 0.5).float()
        out.append(t)
    return torch.stack(out)


def load_single_image(fp, device, mode="RGB"):
    """inputs.py:43-48."""
    import PIL.Image
    if not fp or not device:
        raise ValueError("Function requires file pointer as string and device to store tensor to.")
    return _to_tensor(PIL.Image.open(fp).convert(mode)).to(device)


def compute_edges(images_tensor, device):
    """inputs.py:50-69 — Sobel-3 magnitude + 5x5 Gaussian in float64 (OpenCV REFLECT_101 borders), on the GPU."""
    imgs = images_tensor.detach().to(device=device, dtype=torch.float32).contiguous()
    return _engine_for(device).compute_edges(imgs)


def erode_images(images_tensor, device, kernel=(5, 5)):
    """inputs.py:71-85 — rectangular erosion; OpenCV's default border for erode ignores out-of-image pixels."""
    x = images_tensor.detach().to(device=device, dtype=torch.float32)
    kh, kw = kernel
    return -F.max_pool2d(-x, kernel_size=(kh, kw), stride=1, padding=(kh // 2, kw // 2))


def normalize_homography(hom, height, width):
    """kornia.geometry.conversions.normalize_homography(H, (height,width), (height,width)) restated: pixel-space
    homography -> [-1,1]-normalised one, N·H·N⁻¹ with N = [[2/(w-1),0,-1],[0,2/(h-1),-1],[0,0,1]]."""
    n = torch.tensor([[2.0 / max(width - 1, 1e-14), 0, -1], [0, 2.0 / max(height - 1, 1e-14), -1], [0, 0, 1]],
                     dtype=hom.dtype, device=hom.device)
    return n @ (hom @ torch.linalg.inv(n))


def load_homography(fps, width, height, device, append_zero=True):
    """inputs.py:87-105 — text .mat files -> [B,3,3]; the reference passes (width,height) where kornia expects
    (height,width), which is reproduced here so Homography_Error reads the same."""
    if not fps:
        return None
    if not isinstance(fps, list):
        raise TypeError("Function requires a list of input file paths!")
    mats = [torch.eye(3, dtype=torch.float32)] if append_zero else []
    mats += [torch.tensor(np.loadtxt(fp), dtype=torch.float32) for fp in fps]
    gt = torch.stack(mats).to(device)
    return normalize_homography(gt, height=width, width=height)


def prepare_images(opt, fps_images=None, fps_masks=None, fp_gt=None, fps_hom=None, edges=True):
    """inputs.py:107-127."""
    out = edict()
    out.gt = load_single_image(fp_gt, opt.device)
    out.rgb = load_images(fps_images, opt)
    out.gt_hom = load_homography(fps_hom, opt.W, opt.H, opt.device)
    out.masks = load_images(fps_masks, opt, mode="L", invert_gray=True)
    out.masks_eroded = erode_images(out.masks, opt.device, kernel=(5, 5)) if out.masks is not None else None
    out.gray = load_images(fps_images, opt, mode="L")
    out.edges = compute_edges(out.gray, opt.device) if edges else None
    return out


def move_to_device(x, device):
    """util.py:81-95."""
    if isinstance(x, dict):
        for k, v in x.items():
            x[k] = move_to_device(v, device)
        return x
    if isinstance(x, list):
        return [move_to_device(e, device) for e in x]
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return type(x)(**move_to_device(x._asdict(), device))
    if isinstance(x, torch.Tensor):
        return x.to(device=device)
    return x
