"""Import the UNMODIFIED reference from /root/reference for golden-vector generation.

TEST INFRASTRUCTURE ONLY (see oracle/planar_oracle.py).  Works only in the build container
(`/root/reference` does not exist on the GPU box); nothing in tests/, smoke() or bench.py
imports this module at run time — it is used by `oracle/gen_golden.py` alone.

The reference imports seven pip packages that are absent here (easydict, termcolor, ipdb,
imageio, visdom, kornia, matplotlib — SURVEY.md §8c).  They are stubbed with the minimal real
behaviour the hot path touches; no reference source is edited or copied.  The reference's
`--cpu` bug in `inputs.compute_edges` (`.numpy()` on a grad-requiring CPU tensor,
inputs.py:57-59) is worked around by wrapping that one function so it detaches first.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MARF_REFERENCE_ROOT", "/root/reference")


class _EasyDict(dict):
    """Attribute dict with recursive conversion — the behaviour of easydict.EasyDict the reference uses."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, (list, tuple)):
            value = type(value)(self.__class__(x) if isinstance(x, dict) else x for x in value)
        elif isinstance(value, dict) and not isinstance(value, _EasyDict):
            value = _EasyDict(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k, v in d.items():
            setattr(self, k, v)

    def pop(self, k, *args):
        if hasattr(self, k):
            delattr(self, k)
        return super().pop(k, *args)


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "easydict" not in sys.modules:
        mod("easydict", EasyDict=_EasyDict)
    if "termcolor" not in sys.modules:
        mod("termcolor", colored=lambda s, *a, **k: s)
    for name in ("ipdb", "visdom"):
        if name not in sys.modules:
            mod(name)
    if "imageio" not in sys.modules:
        def imsave(path, arr):
            import PIL.Image
            import numpy as np
            PIL.Image.fromarray(np.asarray(arr)).save(path)
        mod("imageio", imsave=imsave)
    if "kornia" not in sys.modules:
        def normalize_homography(dst_pix_trans_src_pix, dsize_src, dsize_dst):
            """kornia.geometry.conversions.normalize_homography restated from kornia's documentation (kornia is unpinned in
            the reference's requirements.yaml:27 and absent here): N_dst . H . N_src^-1 with
            N(h, w) = [[2/(w-1), 0, -1], [0, 2/(h-1), -1], [0, 0, 1]] (normal_transform_pixel, eps 1e-14 for size-1 axes)."""
            import torch

            def normal_transform_pixel(height, width, like):
                n = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=like.dtype, device=like.device)
                n[0, 0] = n[0, 0] * 2.0 / (width - 1 if width != 1 else 1e-14)
                n[1, 1] = n[1, 1] * 2.0 / (height - 1 if height != 1 else 1e-14)
                return n.unsqueeze(0)
            src_h, src_w = dsize_src
            dst_h, dst_w = dsize_dst
            src_norm = normal_transform_pixel(src_h, src_w, dst_pix_trans_src_pix)
            dst_norm = normal_transform_pixel(dst_h, dst_w, dst_pix_trans_src_pix)
            return dst_norm @ (dst_pix_trans_src_pix @ torch.linalg.inv(src_norm))
        conv = mod("kornia.geometry.conversions", normalize_homography=normalize_homography)
        geo = mod("kornia.geometry", conversions=conv)
        mod("kornia", geometry=geo)
    if "matplotlib" not in sys.modules:
        plt = mod("matplotlib.pyplot")
        mod("matplotlib", pyplot=plt)


def load_reference():
    """Returns (planar_module, warp_module, inputs_module, easydict_class)."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree {REFERENCE_ROOT} not present (only in the build container)")
    _install_stubs()
    for name in ("model", "model.planar", "warp", "inputs", "util", "util_vis", "options"):
        if name in sys.modules and REFERENCE_ROOT not in (getattr(sys.modules[name], "__file__", "") or ""):
            del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib
        planar = importlib.import_module("model.planar")
        warp = importlib.import_module("warp")
        inputs = importlib.import_module("inputs")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    # --cpu workaround for inputs.py:57-59 (the reference calls .numpy() on a tensor that requires grad)
    if not getattr(inputs, "_marf_detach_patch", False):
        orig = inputs.compute_edges

        def compute_edges_detached(images_tensor, device):
            return orig(images_tensor.detach(), device)
        inputs.compute_edges = compute_edges_detached
        inputs._marf_detach_patch = True
    return planar, warp, inputs, sys.modules["easydict"].EasyDict


def make_opt(edict, **over):
    """An `opt` carrying exactly the keys the reference's Graph/Warp/NeuralImageFunction read."""
    base = dict(
        H=360, W=480, patch_H=180, patch_W=240, batch_size=5, device="cpu",
        use_masks=True, use_implicit_mask=False, N_vocab=1500, build_single_masks=False,
        use_edges=False, alpha_initial=0.0, alpha_final=1.0, use_cropped_images=True,
        use_homographies=False, max_iter=3000, barf_c2f=None,
        arch=dict(layers=[None, 256, 256, 256, 256, 3], skip=[], posenc=dict(L_2D=8)),
        warp=dict(type="homography", dof=8, noise_h=0.1, noise_t=0.2, fix_first=True),
        loss_weight=dict(render=0, rgb=0, edge=0, mask=0),
        optim=dict(lr=1e-3, lr_warp=1e-3, lr_mask=1e-3, algo="Adam", sched={}),
        output_path="/tmp/marf_ref_out", tb=None, freq=dict(scalar=20, vis=100), max_epoch=1000,
        dataset="none", seed=0,
    )
    def merge(a, b):
        for k, v in b.items():
            if isinstance(v, dict) and isinstance(a.get(k), dict):
                merge(a[k], v)
            else:
                a[k] = v
    merge(base, over)
    return edict(base)
