"""Device-side optimizer step for the planar plugin: torch.optim.Adam's update for every parameter
of the step in ONE kernel launch (marf_adam_step), including the reference's `warp.fix_first` reset of the loop tail
(model/planar.py:157-158).  Same hyper-parameter surface as the reference's setup_optimizer (model/planar.py:86-99):
per-group learning rates lr / lr_warp / lr_mask, Adam defaults beta=(0.9,0.999), eps=1e-8."""
import ctypes as C

import torch

from . import _lib as L


class FusedAdam:
    def __init__(self, engine, params, grads, lrs, betas=(0.9, 0.999), eps=1e-8, zero_tensor=-1, zero_count=0):
        """params / grads: lists of contiguous fp32 CUDA tensors (grads = views of the flat gradient buffer the
        fused step writes); lrs: per-tensor learning rates."""
        assert len(params) == len(grads) == len(lrs)
        self.engine = engine
        self.params, self.grads = list(params), list(grads)
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.step_count = 0
        n = len(self.params)
        self._io = L.MarfAdamIO()
        self._keep = dict(
            p=L.ptr_array(self.params), g=L.ptr_array(self.grads), m=L.ptr_array(self.exp_avg), v=L.ptr_array(self.exp_avg_sq),
            numel=(C.c_int64 * n)(*[p.numel() for p in self.params]), lr=(C.c_float * n)(*[float(x) for x in lrs]))
        io = self._io
        io.n_tensors = n
        io.params, io.grads, io.exp_avg, io.exp_avg_sq = self._keep["p"], self._keep["g"], self._keep["m"], self._keep["v"]
        io.numel, io.lr = self._keep["numel"], self._keep["lr"]
        io.beta1, io.beta2, io.eps = float(betas[0]), float(betas[1]), float(eps)
        io.zero_tensor, io.zero_count = int(zero_tensor), int(zero_count)

    def zero_grad(self):
        """Gradients are overwritten by every fused step; nothing to clear."""

    def step(self):
        self.step_count += 1
        self._io.step = self.step_count
        e = self.engine
        L.check(e.lib, e.handle, e.lib.marf_adam_step(e.handle, C.byref(self._io), e._stream()), "marf_adam_step")
