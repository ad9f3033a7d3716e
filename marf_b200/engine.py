"""Python host side of the C ABI: one `PlanarEngine` per rank/device.

PyTorch is plumbing here (device memory, streams, torch.distributed); all arithmetic of the step runs in
libmarf_b200.so.  There is no fallback path: constructing an engine without the library or without a B200 raises.
"""
import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib as L


def shard_plan(batch_global: int, h: int, rank: int, world: int, whole_patches: bool = False):
    """whole patches per rank when they divide evenly, else an equal row range of every patch.
    `whole_patches` (the edge term needs whole patches: its Sobel/Gauss stencils read 3 rows of halo): patches are dealt out as
    evenly as they go instead, the first `batch_global % world` ranks take one more, and a rank may end up with none (batch 0).
    Returns (batch, patch_offset, rows, row_offset)."""
    if world <= 1:
        return batch_global, 0, h, 0
    if batch_global % world == 0:
        per = batch_global // world
        return per, rank * per, h, 0
    if whole_patches:
        base, extra = divmod(batch_global, world)
        return base + (1 if rank < extra else 0), rank * base + min(rank, extra), h, 0
    base, extra = divmod(h, world)
    rows = base + (1 if rank < extra else 0)
    row_offset = rank * base + min(rank, extra)
    if rows == 0:
        raise ValueError(f"cannot shard {h} rows over {world} ranks")
    return batch_global, 0, rows, row_offset


class PlanarEngine:
    """Owns a marf_handle plus the flat gradient / loss-sum buffers the step writes."""

    def __init__(self, *, H, W, patch_H, patch_W, batch_size, layers: Sequence[int], skip: Sequence[int] = (),
                 L_2D: Optional[int] = 8, barf_c2f=None, mask_mode=L.MASK_NONE, use_edges=False, use_cropped=True,
                 mask_layers: Sequence[int] = (256, 256, 256, 256, 1), mask_uv_freqs=10, mask_embed_dim=128,
                 edge_label_channels=1, precision="fp32", device=None, rank=0, world=1, max_chunk_pixels=0, n_vocab=1500):
        if not torch.cuda.is_available():
            raise RuntimeError("marf_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = L.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        h = patch_H if use_cropped else H
        w = patch_W if use_cropped else W
        self.h, self.w = h, w
        self.batch_global = batch_size
        self.batch, self.patch_offset, self.rows, self.row_offset = shard_plan(batch_size, h, rank, world, whole_patches=bool(use_edges))
        # a rank that was dealt no patch (more ranks than patches, whole-patch sharding) still needs a handle for the replicated
        # optimizer step and the renders: it evaluates patch 0 and its contributions are zeroed before the exchange (Graph.forward)
        self.idle = self.batch == 0
        if self.idle:
            self.batch, self.patch_offset = 1, 0
        self.rank, self.world = rank, world
        self.mask_mode = mask_mode
        self.use_edges = bool(use_edges)
        self.layers = [int(x) for x in layers]
        self.mask_layers = [int(x) for x in mask_layers]
        self.n_local = self.batch * self.rows * w
        self.n_global = batch_size * h * w
        cfg = L.MarfConfig()
        cfg.abi_version = L.MARF_ABI_VERSION
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg.precision = {"fp32": L.FP32, "bf16": L.BF16}[precision]
        self.precision = precision
        cfg.H, cfg.W, cfg.patch_H, cfg.patch_W, cfg.use_cropped = H, W, patch_H, patch_W, int(bool(use_cropped))
        cfg.batch_global, cfg.batch, cfg.patch_offset = batch_size, self.batch, self.patch_offset
        cfg.rows, cfg.row_offset = self.rows, self.row_offset
        cfg.L = int(L_2D) if L_2D else 0
        cfg.n_layers = len(self.layers)
        for i, k in enumerate(self.layers):
            cfg.layer_out[i] = k
        cfg.skip_mask = sum(1 << int(s) for s in skip)
        cfg.c2f_enabled = int(barf_c2f is not None)
        if barf_c2f is not None:
            cfg.c2f_start, cfg.c2f_end = float(barf_c2f[0]), float(barf_c2f[1])
        cfg.mask_mode = mask_mode
        if mask_mode == L.MASK_IMPLICIT:
            cfg.mask_n_layers = len(self.mask_layers)
            for i, k in enumerate(self.mask_layers):
                cfg.mask_layer_out[i] = k
            cfg.mask_uv_freqs, cfg.mask_embed_dim = mask_uv_freqs, mask_embed_dim
        cfg.use_edges = int(self.use_edges)
        cfg.edge_label_channels = edge_label_channels
        cfg.max_chunk_pixels = int(max_chunk_pixels)
        cfg.mask_n_vocab = int(n_vocab)
        self.cfg = cfg
        handle = C.c_void_p()
        rc = self.lib.marf_create(C.byref(cfg), C.byref(handle))
        if rc != 0:
            msg = self.lib.marf_last_error(None)
            raise L.MarfError(f"marf_create failed (code {rc}): {msg.decode() if msg else '?'}")
        self.handle = handle
        self.sums = torch.zeros(L.N_SUMS, dtype=torch.float64, device=self.device)
        self._io = L.MarfStepIO()
        self._keep = {}
        self.data_version = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.marf_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    @property
    def launches(self) -> int:
        return int(self.lib.marf_launch_count(self.handle))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.marf_workspace_bytes(self.handle))

    PROF_CLASSES = ("k_tc_chain<fwd>", "k_tc_chain<dx>", "k_tc_dw", "k_tc_bwd", "k_tc_gemm<64,warp_grad>")

    def profile(self, enable: bool):
        """Event pairs around the tensor-core launches of the bf16 path (marf_profile); see profile_read."""
        L.check(self.lib, self.handle, self.lib.marf_profile(self.handle, 1 if enable else 0), "marf_profile")

    def profile_read(self):
        """-> {kernel class: (milliseconds summed over the recorded launches, launches)}; clears the record."""
        n = len(self.PROF_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        L.check(self.lib, self.handle, self.lib.marf_profile_read(self.handle, ms, cnt, n), "marf_profile_read")
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.PROF_CLASSES)}

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _chk(t, dtype, name):
        if t is None:
            return
        if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous CUDA {dtype} tensor")

    def _ptrs(self, key, tensors):
        if tensors is None:
            return None
        arr = L.ptr_array(list(tensors))
        self._keep[key] = arr
        return arr

    def fill_io(self, *, mlp_w, mlp_b, warp, rgb, masks=None, masks_eroded=None, edges=None, mask_w=None, mask_b=None,
                embed=None, g_mlp_w=None, g_mlp_b=None, g_warp=None, g_mask_w=None, g_mask_b=None, rgb_pred=None,
                mask_pred=None, edge_pred=None, progress=0.0, coef=(2.0, 0.0, 0.0), norm_rgb=0.0, norm_edge=0.0):
        f32, f64 = torch.float32, torch.float64
        for name, t in [("warp", warp), ("rgb", rgb), ("masks", masks), ("masks_eroded", masks_eroded), ("embed", embed),
                        ("g_warp", g_warp), ("rgb_pred", rgb_pred), ("mask_pred", mask_pred)]:
            self._chk(t, f32, name)
        for name, t in [("edges", edges), ("edge_pred", edge_pred)]:
            self._chk(t, f64, name)
        for group in (mlp_w, mlp_b, mask_w, mask_b, g_mlp_w, g_mlp_b, g_mask_w, g_mask_b):
            if group is not None:
                for t in group:
                    self._chk(t, f32, "parameter/gradient")
        io = self._io
        io.mlp_w = self._ptrs("mlp_w", mlp_w)
        io.mlp_b = self._ptrs("mlp_b", mlp_b)
        io.warp = warp.data_ptr()
        io.mask_w = self._ptrs("mask_w", mask_w)
        io.mask_b = self._ptrs("mask_b", mask_b)
        io.embed = embed.data_ptr() if embed is not None else None
        io.rgb = rgb.data_ptr()
        io.masks = masks.data_ptr() if masks is not None else None
        io.masks_eroded = masks_eroded.data_ptr() if masks_eroded is not None else None
        io.edges = edges.data_ptr() if edges is not None else None
        io.data_version = self.data_version
        io.progress = float(progress)
        io.c_rgb, io.c_mask, io.c_edge = [float(x) for x in coef]
        io.norm_rgb, io.norm_edge = float(norm_rgb), float(norm_edge)
        io.g_mlp_w = self._ptrs("g_mlp_w", g_mlp_w)
        io.g_mlp_b = self._ptrs("g_mlp_b", g_mlp_b)
        io.g_warp = g_warp.data_ptr() if g_warp is not None else None
        io.g_mask_w = self._ptrs("g_mask_w", g_mask_w)
        io.g_mask_b = self._ptrs("g_mask_b", g_mask_b)
        io.rgb_pred = rgb_pred.data_ptr() if rgb_pred is not None else None
        io.mask_pred = mask_pred.data_ptr() if mask_pred is not None else None
        io.edge_pred = edge_pred.data_ptr() if edge_pred is not None else None
        io.loss_sums = self.sums.data_ptr()
        # keep tensors alive until the next fill (the call is asynchronous)
        self._keep["tensors"] = (mlp_w, mlp_b, warp, rgb, masks, masks_eroded, edges, mask_w, mask_b, embed, g_mlp_w,
                                 g_mlp_b, g_warp, g_mask_w, g_mask_b, rgb_pred, mask_pred, edge_pred)
        return io

    def bump_data_version(self):
        """Call when the CONTENTS of rgb / masks / embed change (cached derived inputs are rebuilt)."""
        self.data_version += 1

    # ------------------------------------------------------------------ entry points
    def step(self, **kw):
        io = self.fill_io(**kw)
        L.check(self.lib, self.handle, self.lib.marf_step(self.handle, C.byref(io), self._stream()), "marf_step")
        return self.sums

    def step_forward(self, **kw):
        io = self.fill_io(**kw)
        L.check(self.lib, self.handle, self.lib.marf_step_forward(self.handle, C.byref(io), self._stream()),
                "marf_step_forward")
        return self.sums

    def step_backward(self):
        L.check(self.lib, self.handle, self.lib.marf_step_backward(self.handle, C.byref(self._io), self._stream()),
                "marf_step_backward")
        return self.sums

    def render(self, mlp_w: List[torch.Tensor], mlp_b: List[torch.Tensor], *, crop=False, warp=None, n_patches=1,
               progress=0.0) -> torch.Tensor:
        cfg = self.cfg
        P = (cfg.patch_H * cfg.patch_W) if crop else (cfg.H * cfg.W)
        out = torch.empty(n_patches, P, 3, dtype=torch.float32, device=self.device)
        io = L.MarfRenderIO()
        aw, ab = L.ptr_array(list(mlp_w)), L.ptr_array(list(mlp_b))
        io.mlp_w, io.mlp_b = aw, ab
        io.warp = warp.data_ptr() if warp is not None else None
        io.n_patches, io.crop, io.progress = n_patches, int(bool(crop)), float(progress)
        io.rgb = out.data_ptr()
        L.check(self.lib, self.handle, self.lib.marf_render(self.handle, C.byref(io), self._stream()), "marf_render")
        return out

    def forward_points(self, mlp_w: List[torch.Tensor], mlp_b: List[torch.Tensor], xy: torch.Tensor, progress=0.0) -> torch.Tensor:
        """NeuralImageFunction.forward(coord_2d) (model/planar.py:429-449) for explicit, already warped coordinates
        xy [...,2] -> rgb [...,3] through marf_forward_points."""
        shape = tuple(xy.shape[:-1])
        flat = xy.reshape(-1, 2).to(device=self.device, dtype=torch.float32).contiguous()
        n = flat.shape[0]
        out = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        aw, ab = L.ptr_array(list(mlp_w)), L.ptr_array(list(mlp_b))
        L.check(self.lib, self.handle,
                self.lib.marf_forward_points(self.handle, aw, ab, flat.data_ptr(), n, float(progress), out.data_ptr(), self._stream()),
                "marf_forward_points")
        return out.view(*shape, 3)

    def sl3_to_SL3(self, warp: torch.Tensor) -> torch.Tensor:
        self._chk(warp, torch.float32, "warp")
        n = warp.shape[0]
        out = torch.empty(n, 3, 3, dtype=torch.float32, device=self.device)
        L.check(self.lib, self.handle,
                self.lib.marf_sl3_to_SL3(self.handle, warp.data_ptr(), n, out.data_ptr(), self._stream()), "marf_sl3_to_SL3")
        return out

    def warp_corners(self, warp: torch.Tensor) -> torch.Tensor:
        self._chk(warp, torch.float32, "warp")
        n = warp.shape[0]
        out = torch.empty(n, 4, 2, dtype=torch.float32, device=self.device)
        L.check(self.lib, self.handle,
                self.lib.marf_warp_corners(self.handle, warp.data_ptr(), n, out.data_ptr(), self._stream()),
                "marf_warp_corners")
        return out

    def warp_points(self, xy: torch.Tensor, warp: torch.Tensor) -> torch.Tensor:
        """Warp.warp_grid (warp.py:70-81): xy [n,p,2], warp [n,8] -> [n,p,2]."""
        self._chk(xy, torch.float32, "xy")
        self._chk(warp, torch.float32, "warp")
        n, p = xy.shape[0], xy.shape[1]
        out = torch.empty_like(xy)
        L.check(self.lib, self.handle,
                self.lib.marf_warp_points(self.handle, xy.data_ptr(), warp.data_ptr(), n, p, out.data_ptr(), self._stream()),
                "marf_warp_points")
        return out

    def compute_edges(self, images: torch.Tensor) -> torch.Tensor:
        """inputs.compute_edges (inputs.py:50-69) on device: [n,c,h,w] f32 -> float64."""
        self._chk(images, torch.float32, "images")
        n, c, hh, ww = images.shape
        out = torch.empty(n, c, hh, ww, dtype=torch.float64, device=self.device)
        L.check(self.lib, self.handle,
                self.lib.marf_compute_edges(self.handle, images.data_ptr(), n, c, hh, ww, out.data_ptr(), self._stream()),
                "marf_compute_edges")
        return out

    # ------------------------------------------------------------------ loss values from the sums (device, no sync)
    def loss_scalars(self, sums: torch.Tensor, alpha: float, weights4: Sequence[float], out: torch.Tensor):
        """out[5] (float64, device) <- {rgb, mask, edge, render, all} from the step's sums in ONE launch (marf_loss_scalars)."""
        w = (C.c_double * 4)(*[float(x) for x in weights4])
        L.check(self.lib, self.handle,
                self.lib.marf_loss_scalars(self.handle, C.c_void_p(sums.data_ptr()), float(alpha), w, C.c_void_p(out.data_ptr()),
                                           self._stream()), "marf_loss_scalars")
        return out

    def loss_values(self, sums: Optional[torch.Tensor] = None):
        """(rgb, mask, edge) as 0-dim float64 device tensors (model/planar.py:362-370)."""
        s = self.sums if sums is None else sums
        zero = torch.zeros((), dtype=torch.float64, device=s.device)
        rgb = s[L.S_RGB] / s[L.N_RGB]
        mask = s[L.S_MASK] / s[L.N_MASK] if self.mask_mode == L.MASK_IMPLICIT else zero
        edge = s[L.S_EDGE] / s[L.N_EDGE] if self.use_edges else zero
        return rgb, mask, edge
