"""Host-side mirror of the reference's geometry helpers (`warp.py`): same class and method names, same argument
meaning.  The arithmetic runs in libmarf_b200.so (marf_sl3_to_SL3 / marf_warp_points / marf_warp_corners); inside
the training step the grid and the warp are computed analytically in the fused kernels and none of this is called.
"""
import torch

from . import _lib as L
from .engine import PlanarEngine

_util_engines = {}


def _util_engine(device, opt=None) -> PlanarEngine:
    """A minimal handle for the stateless geometry entry points."""
    key = (str(device), None if opt is None else (opt.H, opt.W, opt.patch_H, opt.patch_W, opt.batch_size))
    if key not in _util_engines:
        if opt is None:
            kw = dict(H=4, W=4, patch_H=2, patch_W=2, batch_size=1)
        else:
            kw = dict(H=opt.H, W=opt.W, patch_H=opt.patch_H, patch_W=opt.patch_W, batch_size=opt.batch_size)
        _util_engines[key] = PlanarEngine(layers=[3], L_2D=None, mask_mode=L.MASK_NONE, device=device,
                                          max_chunk_pixels=128, **kw)
    return _util_engines[key]


class Warp:
    """warp.py:5-93."""

    def __init__(self, opt):
        self.opt = opt
        self.max_h, self.max_w = opt.H, opt.W
        self.crop_h, self.crop_w = opt.patch_H, opt.patch_W
        self.y_crop = (opt.H // 2 - opt.patch_H // 2, opt.H // 2 + opt.patch_H // 2)
        self.x_crop = (opt.W // 2 - opt.patch_W // 2, opt.W // 2 + opt.patch_W // 2)
        longest = max(opt.H, opt.W)
        self.norm_h, self.norm_w = opt.H / longest, opt.W / longest
        self.batch_size = opt.batch_size
        self.device = opt.device
        self.warp_type = opt.warp.type
        self.dof = opt.warp.dof

    def to_hom(self, matrix):
        return torch.cat([matrix, torch.ones_like(matrix[..., :1])], dim=-1)

    def _axis(self, lo, hi, n, norm):
        i = torch.arange(lo, hi, dtype=torch.float32, device=self.device)
        return ((i + 0.5) / n * 2 - 1) * norm

    def get_normalized_pixel_grid(self, crop=False):
        """[B, h*w, 2] in (x, y) order, row-major over (row, col); warp.py:33-68.  Setup/diagnostic use only —
        the step derives the same coordinates from the pixel index on device."""
        ys = self._axis(*(self.y_crop if crop else (0, self.max_h)), self.max_h, self.norm_h)
        xs = self._axis(*(self.x_crop if crop else (0, self.max_w)), self.max_w, self.norm_w)
        Y, X = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([X, Y], dim=-1).view(-1, 2).repeat(self.batch_size, 1, 1)

    def warp_grid(self, xy_grid, warp):
        """warp.py:70-81: homography apply with H = expm(A(warp))."""
        if self.warp_type != "homography" or self.dof != 8:
            raise AssertionError("only warp.type=homography, dof=8 is defined (warp.py:72-80)")
        eng = _util_engine(xy_grid.device)
        return eng.warp_points(xy_grid.contiguous().float(), warp.detach().contiguous().float())

    def warp_corners(self, warp_param):
        """warp.py:83-93: the four crop corners under each patch's warp, [B,4,2]."""
        eng = _util_engine(warp_param.device, self.opt)
        return eng.warp_corners(warp_param.detach().contiguous().float())


class Lie:
    """warp.py:95-108."""

    def sl3_to_SL3(self, h):
        eng = _util_engine(h.device)
        return eng.sl3_to_SL3(h.detach().contiguous().float())


lie = Lie()
