"""Host-side mirror of the reference's planar plugin (`model/planar.py`): the same `Model` / `Graph` /
`NeuralImageFunction` / `ImplicitMask` / `PosEmbedding` classes, method names and `var` / `loss` contracts,
with the step itself (Graph.forward + compute_loss + backward, model/planar.py:329-391) executed by
libmarf_b200.so through the C ABI.  PyTorch holds the parameters, the optimizer and the NCCL plumbing.

Deviations from the reference, all deliberate and documented in DESIGN.md:
  * `loss.render` carries the autograd edge of the fused step (its backward hands out d(loss.all)/dθ computed
    in-kernel); `loss.rgb/.mask/.edge` are detached scalars.  `loss.all.backward()` therefore works unchanged.
  * `embedding_view.weight.grad` is not produced (the reference computes it but never applies it:
    the embedding is not in the optimizer, model/planar.py:89-96).
  * `build_single_masks` (per-image mask nets forced onto the CPU, model/planar.py:322-324,347) is refused.
  * under torchrun the patches (or rows) are sharded over ranks and gradients are all-reduced.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from . import inputs
from .attrdict import AttrDict as edict
from .engine import PlanarEngine
from .warp import Lie, Warp


def get_layer_dims(layers):
    """(k_in, k_out) pairs of consecutive entries (util.py:105-108)."""
    return list(zip(layers[:-1], layers[1:]))


def _dist_on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


# ============================ networks (parameter containers) ============================

class NeuralImageFunction(torch.nn.Module):
    """model/planar.py:395-471 — parameters live here; arithmetic runs in the library."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.define_network()
        self.progress = torch.nn.Parameter(torch.tensor(0.))
        self._engine = None

    def define_network(self):
        """nn.Linear stack in the reference's RNG order, with the c2f first-layer rescale (model/planar.py:410-427)."""
        opt = self.opt
        d_in = 2 + 4 * opt.arch.posenc.L_2D if opt.arch.posenc else 2
        self.input_2d_dim = d_in
        self.mlp = torch.nn.ModuleList()
        for li, (k_in, k_out) in enumerate(get_layer_dims(opt.arch.layers)):
            k_in = d_in if li == 0 else k_in
            if li in opt.arch.skip:
                k_in += d_in
            layer = torch.nn.Linear(k_in, k_out)
            if opt.barf_c2f and li == 0:
                scale = np.sqrt(d_in / 2.)
                layer.weight.data *= scale
                layer.bias.data *= scale
            self.mlp.append(layer)

    def weights(self):
        return [l.weight for l in self.mlp], [l.bias for l in self.mlp]

    @torch.no_grad()
    def forward(self, coord_2d=None, *, crop=False, warp=None, n_patches=1):
        """model/planar.py:429-449.  With `coord_2d` [...,2] (the reference's signature: already warped normalised
        coordinates) -> rgb [...,3] through marf_forward_points.  Without it the grid is generated on device (marf_render):
        `crop` selects get_normalized_pixel_grid(crop=...), `warp` optional per-patch sl(3) parameters."""
        ws, bs = self.weights()
        if coord_2d is not None:
            return self._engine.forward_points([w.detach() for w in ws], [b.detach() for b in bs], coord_2d,
                                               progress=float(self.progress))
        return self._engine.render([w.detach() for w in ws], [b.detach() for b in bs], crop=crop, warp=warp,
                                   n_patches=n_patches, progress=float(self.progress))


class ImplicitMask(torch.nn.Module):
    """model/planar.py:475-488 — 426→256→256→256→256→1 parameter container."""

    def __init__(self, latent=3 * 128, W=256, in_channels_dir=42):
        super().__init__()
        nn = torch.nn
        self.mask_mapping = nn.Sequential(
            nn.Linear(latent + in_channels_dir, W), nn.ReLU(True),
            nn.Linear(W, W), nn.ReLU(True),
            nn.Linear(W, W), nn.ReLU(True),
            nn.Linear(W, W), nn.ReLU(True),
            nn.Linear(W, 1), nn.Sigmoid())

    def linears(self):
        return [m for m in self.mask_mapping if isinstance(m, torch.nn.Linear)]


class PosEmbedding(torch.nn.Module):
    """model/planar.py:491-518 — only the frequency table is kept; the embedding is computed in-kernel."""

    def __init__(self, max_logscale, N_freqs, logscale=True):
        super().__init__()
        self.N_freqs = N_freqs
        self.freqs = 2 ** torch.linspace(0, max_logscale, N_freqs) if logscale else torch.linspace(1, 2 ** max_logscale, N_freqs)
        if not logscale or max_logscale != N_freqs - 1:
            raise NotImplementedError("the fused kernel implements freqs 2^0..2^(N-1) (PosEmbedding(N-1, N))")


class _FusedStepGrad(torch.autograd.Function):
    """Autograd edge of the fused step: forward passes the render-loss value through; backward returns the
    gradients of loss.all the library already computed, scaled by the upstream gradient."""

    @staticmethod
    def forward(ctx, value, graph, inv_render_weight, *params):
        ctx.graph = graph
        ctx.inv = inv_render_weight
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        graph = ctx.graph
        scaled = graph._grad_flat * (g * ctx.inv).to(torch.float32)
        return (None, None, None) + tuple(scaled[a:b].view(shape) for a, b, shape in graph._grad_slices)


# ============================ computation graph ============================

class Graph(torch.nn.Module):
    """model/planar.py:296-391."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.batch_size = opt.batch_size
        self.neural_image = NeuralImageFunction(opt)
        self.warp = Warp(opt)
        self.warp_param = torch.nn.Embedding(self.batch_size, opt.warp.dof)
        torch.nn.init.zeros_(self.warp_param.weight)
        self.h = opt.patch_H if opt.use_cropped_images else opt.H
        self.w = opt.patch_W if opt.use_cropped_images else opt.W
        self.max_iter = opt.max_iter
        self.it = 0
        if opt.use_implicit_mask:
            if opt.build_single_masks:
                raise NotImplementedError("build_single_masks (CPU-resident per-image mask nets) is not supported")
            self.embedding_uv = PosEmbedding(10 - 1, 10)
            self.implicit_mask = ImplicitMask()
            self.embedding_view = torch.nn.Embedding(opt.N_vocab, 128)
        self.engine = None
        self._grad_flat = None
        self._grad_slices = None
        self._local = None
        self._norms = (0.0, 0.0)
        self.data_parallel = None        # None: follow torch.distributed; False: force single-rank behaviour

    # ------------------------------------------------------------------ engine plumbing
    def _mask_mode(self):
        if self.opt.use_implicit_mask:
            return L.MASK_IMPLICIT
        return L.MASK_DISK if self.opt.use_masks else L.MASK_NONE

    def step_params(self):
        """Parameters the step differentiates, in the order of the flat gradient buffer."""
        ws, bs = self.neural_image.weights()
        ps = list(ws) + list(bs) + [self.warp_param.weight]
        if self.opt.use_implicit_mask:
            lins = self.implicit_mask.linears()
            ps += [l.weight for l in lins] + [l.bias for l in lins]
        return ps

    def _dp(self):
        return _dist_on() if self.data_parallel is None else bool(self.data_parallel)

    def _ensure_engine(self):
        if self.engine is not None:
            return
        opt = self.opt
        dev = self.warp_param.weight.device
        if dev.type != "cuda":
            raise RuntimeError("Graph must live on a CUDA device (marf_b200 has no CPU path)")
        rank = dist.get_rank() if self._dp() else 0
        world = dist.get_world_size() if self._dp() else 1
        self.engine = PlanarEngine(
            H=opt.H, W=opt.W, patch_H=opt.patch_H, patch_W=opt.patch_W, batch_size=opt.batch_size,
            layers=list(opt.arch.layers[1:]), skip=list(opt.arch.skip or []),
            L_2D=opt.arch.posenc.L_2D if opt.arch.posenc else None,
            barf_c2f=tuple(opt.barf_c2f) if opt.barf_c2f else None, mask_mode=self._mask_mode(),
            use_edges=bool(opt.use_edges), use_cropped=bool(opt.use_cropped_images),
            precision=opt.get("precision", "fp32") or "fp32", device=dev, rank=rank, world=world,
            max_chunk_pixels=int(opt.get("max_chunk_pixels", 0) or 0), n_vocab=int(opt.get("N_vocab", 1500) or 1500))
        self.neural_image._engine = self.engine
        ps = self.step_params()
        total = sum(p.numel() for p in ps)
        self._grad_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._grad_slices, off = [], 0
        for p in ps:
            self._grad_slices.append((off, off + p.numel(), tuple(p.shape)))
            off += p.numel()
        self._grad_views = [self._grad_flat[a:b].view(shape) for a, b, shape in self._grad_slices]
        # data parallel: the step's two small exchanges as one-shot all-reduces over NVLink peer memory (marf_peer_allreduce);
        # the engine then writes its gradients and loss sums into this rank's symmetric buffer.  NCCL (dist.all_reduce) when
        # symmetric memory is not available or MARF_NCCL_ALLREDUCE is set.
        self._peer, self._step_grad_views = None, self._grad_views
        # (measured on one 8x B200 box: the one-shot exchange wins at 2 and 4 ranks — 0.935 vs 0.956 ms/step at 2 — and loses at
        #  8, where every rank reads 8 x 2 MB over NVLink while NCCL reduces inside the switch (NVLS): 1.030 vs 0.990 ms/step)
        #  -> gradients over peer memory up to 4 ranks; the 64-byte loss sums (pure latency) over peer memory at any size)
        peer_max = int(os.environ.get("MARF_PEER_ALLREDUCE_MAX_WORLD", "4"))
        self._peer_grads = False
        if self._dp():
            # replicas must start identical whatever each rank's RNG did (e.g. --seed= gives every rank its own seed)
            for prm in self.parameters():
                dist.broadcast(prm.data, src=0)
            # the peer path needs every rank on one NVLink box, and every rank must take the same path: agree collectively
            one_box = int(os.environ.get("LOCAL_WORLD_SIZE", world)) == world
            want = 1 < world <= 8 and one_box and not os.environ.get("MARF_NCCL_ALLREDUCE")
            peer = None
            if want:
                try:
                    from .peer import PeerAllReduce
                    peer = PeerAllReduce(dev, total)
                except Exception as ex:            # pylint: disable=broad-except
                    print(f"[marf_b200] rank {rank}: peer all-reduce unavailable ({ex!r})", flush=True)
            ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                self._peer = peer
                self._peer_grads = world <= peer_max
                if self._peer_grads:
                    self._step_grad_views = [self._peer.grad_local[a:b].view(shape) for a, b, shape in self._grad_slices]
                self.engine.sums = self._peer.sums
            elif want and rank == 0:
                print("[marf_b200] peer all-reduce not available on every rank; all ranks use NCCL", flush=True)
        e = self.engine
        self._rgb_pred = torch.zeros(e.batch, e.rows * e.w, 3, dtype=torch.float32, device=dev)
        self._mask_pred = torch.zeros(e.batch, e.rows * e.w, 1, dtype=torch.float32, device=dev) \
            if opt.use_implicit_mask else None
        self._edge_pred = torch.zeros(e.batch, 3, e.rows, e.w, dtype=torch.float64, device=dev) if opt.use_edges else None

    def _local_images(self, images):
        """This rank's shard of the resident inputs, cut once (and again if the container changes)."""
        key = (id(images), images.rgb.data_ptr())
        if self._local is not None and self._local[0] == key:
            return self._local[1]
        e = self.engine

        def cut(t, dtype):
            if t is None:
                return None
            return t[e.patch_offset:e.patch_offset + e.batch, :, e.row_offset:e.row_offset + e.rows].to(
                device=e.device, dtype=dtype).contiguous()

        loc = edict(rgb=cut(images.rgb, torch.float32), masks=cut(images.get("masks"), torch.float32),
                    masks_eroded=cut(images.get("masks_eroded"), torch.float32), edges=cut(images.get("edges"), torch.float64))
        # global loss normalisers that do not depend on the forward pass
        n_rgb = n_edge = 0.0
        if self._dp() and not self.opt.use_implicit_mask:
            if self.opt.use_masks:
                s = torch.stack([loc.masks.double().sum(),
                                 loc.masks_eroded.double().sum() if loc.masks_eroded is not None else torch.zeros((), dtype=torch.float64, device=e.device)])
                if e.idle:                     # (a rank without a patch of its own contributes nothing)
                    s.zero_()
                dist.all_reduce(s)
                n_rgb, n_edge = 3.0 * float(s[0]), 3.0 * float(s[1])
            else:
                n_rgb = n_edge = 3.0 * e.n_global
        self._norms = (n_rgb, n_edge)
        self._local = (key, loc)
        e.bump_data_version()
        return loc

    def _zero_contribution(self, sums, grads):
        """A rank that holds no patch of its own (engine.idle) evaluated patch 0 as a stand-in: drop what it computed."""
        if sums is not None:
            sums.zero_()
        if grads:
            (self._peer.grad_local if self._peer_grads else self._grad_flat).zero_()

    def _allreduce(self, sums, grads):
        """The exchange step of data parallelism: loss sums (in place) and / or the flat gradient buffer."""
        if sums is not None:
            if self._peer is not None:
                self._peer.allreduce_sums()
            else:
                dist.all_reduce(sums)
        if grads:
            if self._peer_grads:
                self._peer.allreduce_grads(self._grad_flat)
            else:
                dist.all_reduce(self._grad_flat)

    def loss_coefficients(self):
        """(c_rgb, c_mask, c_edge, alpha): loss.all = c_rgb·rgb + c_mask·mask + c_edge·edge once `render` is
        expanded (model/planar.py:359,371-378) and weighted by 10**loss_weight (:177-184)."""
        opt = self.opt
        alpha = opt.alpha_initial + (opt.alpha_final - opt.alpha_initial) * (self.it / self.max_iter) if opt.use_edges else 0
        lw = opt.loss_weight

        def p(key):
            return 0.0 if lw.get(key) is None else 10 ** float(lw[key])
        if lw.render is None:
            return 0.0, 0.0, 0.0, alpha
        return p("render") * (1 - alpha) + p("rgb"), p("render") * 0.5 + p("mask"), p("render") * alpha + p("edge"), alpha

    # ------------------------------------------------------------------ the step
    def forward(self, var, mode=None):  # pylint: disable=unused-argument
        """Image and mask predictions for the current warps — and, fused with them, the losses and every
        gradient of loss.all (consumed by compute_loss / loss.all.backward())."""
        self._ensure_engine()
        opt, e = self.opt, self.engine
        loc = self._local_images(var.images)
        ws, bs = self.neural_image.weights()
        implicit = bool(opt.use_implicit_mask)
        nl = len(ws)
        gv = self._step_grad_views
        kw = dict(mlp_w=[w.detach() for w in ws], mlp_b=[b.detach() for b in bs], warp=self.warp_param.weight.detach(),
                  rgb=loc.rgb, masks=loc.masks if (opt.use_masks and not implicit) else None,
                  masks_eroded=loc.masks_eroded if (opt.use_masks and not implicit and opt.use_edges) else None,
                  edges=loc.edges if opt.use_edges else None,
                  g_mlp_w=gv[:nl], g_mlp_b=gv[nl:2 * nl], g_warp=gv[2 * nl],
                  rgb_pred=self._rgb_pred, mask_pred=self._mask_pred, edge_pred=self._edge_pred,
                  progress=float(self.neural_image.progress.data))
        if implicit:
            lins = self.implicit_mask.linears()
            nm = len(lins)
            kw.update(mask_w=[l.weight.detach() for l in lins], mask_b=[l.bias.detach() for l in lins],
                      embed=self.embedding_view.weight.detach(),
                      g_mask_w=gv[2 * nl + 1:2 * nl + 1 + nm], g_mask_b=gv[2 * nl + 1 + nm:2 * nl + 1 + 2 * nm])
        c_rgb, c_mask, c_edge, _ = self.loss_coefficients()
        kw["coef"] = (c_rgb, c_mask, c_edge)
        if not self._dp():
            sums = e.step(**kw)
        elif not implicit:
            kw["norm_rgb"], kw["norm_edge"] = self._norms
            sums = e.step(**kw)
            if e.idle:
                self._zero_contribution(sums, grads=True)
            self._allreduce(sums, grads=True)
        else:
            sums = e.step_forward(**kw)
            if e.idle:
                self._zero_contribution(sums, grads=False)
            self._allreduce(sums, grads=False)
            e.step_backward()
            if e.idle:
                self._zero_contribution(None, grads=True)
            self._allreduce(None, grads=True)
        self._sums = sums
        B = e.batch
        var.rgb_prediction = self._rgb_pred
        var.rgb_prediction_map = self._rgb_pred.view(B, e.rows, e.w, 3).permute(0, 3, 1, 2)
        var.edge_prediction = self._edge_pred
        if implicit:
            var.mask_prediction = self._mask_pred
            var.mask_prediction_map = self._mask_pred.view(B, e.rows, e.w, 1).permute(0, 3, 1, 2)
        return var

    def compute_loss(self, var, mode=None):  # pylint: disable=unused-argument
        """model/planar.py:355-380 — loss dict from the sums the fused step left on the device."""
        loss = edict()
        opt = self.opt
        _, _, _, alpha = self.loss_coefficients()
        if opt.loss_weight.render is not None and not torch.is_grad_enabled():
            # fused-optimizer path (no autograd graph wanted): all five scalars from one launch; `all` is precomputed for
            # Model.summarize_loss (which recognises the marker and does not launch a dozen 0-dim tensor ops)
            lw = opt.loss_weight
            w4 = [0.0 if lw[k] is None else 10 ** float(lw[k]) for k in ("render", "rgb", "mask", "edge")]
            self._loss_out = getattr(self, "_loss_out", None)
            if self._loss_out is None:
                self._loss_out = torch.zeros(2, 5, dtype=torch.float64, device=self._sums.device)
            out = self.engine.loss_scalars(self._sums, alpha, w4, self._loss_out[self.it & 1])   # (double-buffered: the caller
            loss.render, loss.rgb, loss.mask, loss.edge = out[3], out[0], out[1], out[2]          #  may read step i-1's loss late)
            self._fused_all = out[4]
            self.it += 1
            return loss
        if opt.loss_weight.render is not None:
            rgb, mask, edge = self.engine.loss_values(self._sums)
            render = (1 - alpha) * rgb + 0.5 * mask + alpha * edge
            if torch.is_grad_enabled():
                render = _FusedStepGrad.apply(render, self, 1.0 / 10 ** float(opt.loss_weight.render), *self.step_params())
            loss.render = render
            loss.rgb, loss.mask, loss.edge = rgb, mask, edge
        self.it += 1
        return loss

    def mse_loss(self, pred, labels, masks=None):
        """model/planar.py:382-391 (metric use, e.g. Mask_Error; inside the step the loss is computed in-kernel)."""
        if masks is None:
            return ((pred.contiguous() - labels) ** 2).mean()
        return (((pred.contiguous() - labels) * masks) ** 2).sum() / (masks.sum() * 3)


# ============================ the plugin ============================

class Model(torch.nn.Module):
    """model/planar.py:31-292 — the 5-call plugin contract of train.py:23-31."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.batch_size = opt.batch_size
        self.dataset = opt.dataset
        os.makedirs(opt.output_path, exist_ok=True)
        self.warp = Warp(opt)
        self.images = None
        self.graph = None
        self.optim = None
        self.sched = None
        self.tb = None
        self.box_colors = None
        self.vis_path = None
        self.video_fname = None
        self.timer = None
        self.warp_pert = None
        self.ep = self.it = self.vis_it = 0
        self.lie = Lie()
        self.fused_tail = None

    def load_dataset(self):
        """model/planar.py:59-78, or the synthetic scene generator when opt.synthetic.enabled."""
        print("loading dataset...")
        opt = self.opt
        syn = opt.get("synthetic") or {}
        if syn.get("enabled"):
            from . import synth
            self.images = synth.make_scene(opt, seed=int(syn.get("seed", 0)), occluders=bool(syn.get("occluders", False)))
            return
        root = f"data/planar/{self.dataset}"
        B = self.batch_size
        self.images = inputs.prepare_images(
            opt,
            fps_images=[f"{root}/{i}.png" for i in range(B)],
            fps_masks=[f"{root}/{i}-m.png" for i in range(B)] if opt.use_masks else None,
            fp_gt=f"{root}/gt.png",
            fps_hom=[f"{root}/H_0_{i}.mat" for i in range(1, B)] if opt.use_homographies else None,
            edges=True if opt.use_edges else None)

    def build_networks(self):
        print("building networks...")
        self.graph = Graph(self.opt).to(self.opt.device)

    def setup_optimizer(self):
        """model/planar.py:86-104 — one optimizer, param groups (lr, lr_warp, lr_mask), optional scheduler."""
        print("setting up optimizers...")
        opt = self.opt
        groups = [dict(params=self.graph.neural_image.parameters(), lr=opt.optim.lr),
                  dict(params=self.graph.warp_param.parameters(), lr=opt.optim.lr_warp)]
        if opt.use_implicit_mask:
            groups.append(dict(params=self.graph.implicit_mask.parameters(), lr=opt.optim.lr_mask))
        self.optim = getattr(torch.optim, opt.optim.algo)(groups)
        if opt.optim.sched:
            kwargs = {k: v for k, v in opt.optim.sched.items() if k != "type"}
            self.sched = getattr(torch.optim.lr_scheduler, opt.optim.sched.type)(self.optim, **kwargs)
        if opt.get("fused_optimizer"):
            if opt.optim.algo != "Adam" or opt.optim.sched:
                raise NotImplementedError("fused_optimizer implements optim.algo=Adam without a scheduler")
            from .optim import FusedAdam
            g = self.graph
            g._ensure_engine()
            ps = g.step_params()
            n_img = 2 * len(g.neural_image.mlp)
            lrs = [opt.optim.lr] * n_img + [opt.optim.lr_warp] + [opt.optim.lr_mask] * (len(ps) - n_img - 1)
            self.fused_tail = FusedAdam(g.engine, [p.data for p in ps], g._grad_views, lrs, zero_tensor=n_img,
                                        zero_count=opt.warp.dof if opt.warp.fix_first else 0)

    def setup_visualizer(self):
        print("setting up visualizers...")
        if self.opt.tb and self._is_main():
            from torch.utils import tensorboard
            self.tb = tensorboard.SummaryWriter(log_dir=self.opt.output_path, flush_secs=10)
        self.vis_path = f"{self.opt.output_path}/vis"
        os.makedirs(self.vis_path, exist_ok=True)
        self.video_fname = f"{self.opt.output_path}/vis.mp4"

    @staticmethod
    def _is_main():
        return not _dist_on() or dist.get_rank() == 0

    # ------------------------------------------------------------------ checkpoints
    # The reference's config carries `load` / `resume` (options/planar.yaml:31,88) but no code behind them; here:
    #   --resume        restart from <output_path>/model.ckpt if it exists (written every freq.ckpt iterations and at the end)
    #   --load=<file>   initialise the networks and warps from a checkpoint, iteration counter reset
    def checkpoint_path(self):
        return os.path.join(self.opt.output_path, "model.ckpt")

    def save_checkpoint(self, path=None):
        """Networks, warps, optimizer moments and iteration counters; rank 0 writes (the replicas are identical)."""
        if not self._is_main():
            return None
        path = path or self.checkpoint_path()
        ck = dict(version=2, it=self.it, graph_it=self.graph.it, vis_it=self.vis_it, graph=self.graph.state_dict())
        if self.fused_tail is not None:
            ft = self.fused_tail
            ck["fused_adam"] = dict(step=ft.step_count, exp_avg=[t.clone() for t in ft.exp_avg], exp_avg_sq=[t.clone() for t in ft.exp_avg_sq])
        elif self.optim is not None:
            ck["optim"] = self.optim.state_dict()
        tmp = path + ".tmp"
        torch.save(ck, tmp)
        os.replace(tmp, path)                      # (never leaves a half-written checkpoint behind)
        return path

    def load_checkpoint(self, path, resume=True):
        """resume=True: also optimizer state and iteration counters (continue the run); False: parameters only."""
        ck = torch.load(path, map_location=self.opt.device, weights_only=True)     # tensors, ints and dicts only
        self.graph.load_state_dict(ck["graph"])
        # cached derived inputs (mask-head features gathered from embedding_view) belong to the old parameters
        self.graph._local = None
        if self.graph.engine is not None:
            self.graph.engine.bump_data_version()
        if not resume:
            # parameters only: the schedule restarts (the checkpoint's `progress` would open the c2f bands at iteration 0)
            self.graph.neural_image.progress.data.fill_(0.0)
            return 0
        self.it, self.graph.it = int(ck["it"]), int(ck["graph_it"])
        self.vis_it = int(ck.get("vis_it", 0))
        self.graph.neural_image.progress.data.fill_(self.it / self.opt.max_iter)
        if (self.fused_tail is not None) != ("fused_adam" in ck):
            print(f"[marf_b200] WARNING: checkpoint holds {'fused' if 'fused_adam' in ck else 'torch.optim'} Adam state but this run "
                  f"uses {'--fused_optimizer' if self.fused_tail is not None else 'torch.optim'}: the moments restart from zero", flush=True)
        if self.fused_tail is not None and "fused_adam" in ck:
            ft, st = self.fused_tail, ck["fused_adam"]
            ft.step_count = int(st["step"])
            for dst, src in zip(ft.exp_avg, st["exp_avg"]):
                dst.copy_(src)
            for dst, src in zip(ft.exp_avg_sq, st["exp_avg_sq"]):
                dst.copy_(src)
        elif self.optim is not None and "optim" in ck:
            self.optim.load_state_dict(ck["optim"])
        return self.it

    def train(self, mode=True):  # pylint: disable=arguments-differ,unused-argument
        """model/planar.py:136-170."""
        import tqdm
        print("TRAINING START")
        self.timer = edict(start=time.time(), it_mean=None)
        self.graph.train()
        var = edict(idx=torch.arange(self.batch_size))
        var.images = self.images
        var = inputs.move_to_device(var, self.opt.device)
        self.images = var.images
        start = 0
        if self.opt.get("load"):
            self.load_checkpoint(self.opt.load, resume=False)
            print(f"initialised from {self.opt.load}")
        if self.opt.get("resume") and os.path.exists(self.checkpoint_path()):
            start = self.load_checkpoint(self.checkpoint_path(), resume=True)
            print(f"resumed from {self.checkpoint_path()} at iteration {start}")
        freq_ckpt = int(self.opt.freq.get("ckpt", 0) or 0)
        loader = tqdm.trange(start, self.opt.max_iter, desc="Training", leave=False, disable=not self._is_main())
        with torch.no_grad():
            var = self.graph.forward(var)
        if start == 0:
            self.visualize(var, step=0)
        for _ in loader:
            self.train_iteration(var, loader)
            if self.opt.warp.fix_first:
                self.graph.warp_param.weight.data[0] = 0
            if freq_ckpt and self.it % freq_ckpt == 0:
                self.save_checkpoint()
        self.save_checkpoint()
        if self._is_main() and os.system("command -v ffmpeg > /dev/null 2>&1") == 0:
            os.system(f"ffmpeg -y -framerate 30 -i {self.vis_path}/%d.png -pix_fmt yuv420p {self.video_fname}")
        if self.tb:
            self.tb.flush()
            self.tb.close()
        print("TRAINING DONE")

    def summarize_loss(self, loss):
        """model/planar.py:172-185 — Σ 10**w · loss[k].  The Inf/NaN asserts of the reference (8 host syncs per
        step) are replaced by the device-side MARF_NONFINITE counter, read with the scalars every freq.scalar steps."""
        assert "all" not in loss
        fused_all = getattr(self.graph, "_fused_all", None)
        if fused_all is not None and not torch.is_grad_enabled():
            self.graph._fused_all = None
            loss.update(all=fused_all)          # computed by marf_loss_scalars with the same weights (Graph.compute_loss)
            return loss
        total = 0.
        for key in loss:
            assert key in self.opt.loss_weight
            assert loss[key].shape == ()
            if self.opt.loss_weight[key] is not None:
                total = total + 10 ** float(self.opt.loss_weight[key]) * loss[key]
        loss.update(all=total)
        return loss

    def train_iteration(self, var, loader):
        """model/planar.py:187-209."""
        self.timer.it_start = time.time()
        if self.fused_tail is not None:
            # device-side tail: the step already produced d(loss.all)/dθ; one Adam launch follows
            with torch.no_grad():
                var = self.graph.forward(var, mode="train")
                loss = self.summarize_loss(self.graph.compute_loss(var, mode="train"))
            self.fused_tail.step()
        else:
            self.optim.zero_grad()
            var = self.graph.forward(var, mode="train")
            loss = self.graph.compute_loss(var, mode="train")
            loss = self.summarize_loss(loss)
            loss.all.backward()
            self.optim.step()
        if self.sched:
            pass  # the reference constructs the scheduler but never steps it (model/planar.py:101-104)
        if (self.it + 1) % self.opt.freq.scalar == 0:
            self.check_finite()
            if self.tb:
                self.log_scalars(loss, var, step=self.it + 1, split="train")
            if loader is not None and hasattr(loader, "set_postfix"):
                loader.set_postfix(it=self.it + 1, loss=f"{float(loss.all):.3f}")
        if (self.it + 1) % self.opt.freq.vis == 0:
            self.visualize(var, step=self.it + 1, split="train")
        self.it += 1
        self.timer.it_end = time.time()
        self.graph.neural_image.progress.data.fill_(self.it / self.opt.max_iter)
        return loss

    def check_finite(self):
        """The reference's per-step asserts (model/planar.py:181-182) and the IndexError its colour embedding raises for
        indices outside the table (model/planar.py:344), from the counters the step left on the device."""
        sums = self.graph._sums.cpu()
        bad, idx = float(sums[L.NONFINITE]), float(sums[L.BAD_INDEX])
        assert bad == 0.0, f"{int(bad)} non-finite predictions in the last step (loss is Inf/NaN)"
        if idx != 0.0:
            raise IndexError(f"{int(idx)} colour indices trunc(rgb) outside the embedding table (index out of range in self); "
                             "precision=bf16 serves images in [0,2) only")

    @torch.no_grad()
    def predict_entire_image(self):
        """model/planar.py:211-217 — full-canvas render, [3,H,W] on the CPU."""
        self.graph._ensure_engine()
        rgb = self.graph.neural_image.forward(crop=False, n_patches=1)
        return rgb.view(self.opt.H, self.opt.W, 3).detach().cpu().permute(2, 0, 1)

    def homography_error(self, pred_hom, gt_hom):
        """model/planar.py:219-223."""
        pred_h = self.lie.sl3_to_SL3(pred_hom)
        return torch.norm((pred_h - gt_hom) ** 2).mean()

    @torch.no_grad()
    def corner_error_px(self, gt_warp):
        """Mean patch-corner alignment error in pixels against ground-truth sl(3) parameters (synthetic scenes);
        the quantity BASELINE.json's 0.1 px criterion is stated in.  Uses Warp.warp_corners (warp.py:83-93)."""
        pred = self.warp.warp_corners(self.graph.warp_param.weight)
        gt = self.warp.warp_corners(gt_warp.to(pred.device))
        scale = max(self.opt.H, self.opt.W) / 2.0          # normalized units -> pixels (warp.py:38-49)
        return ((pred - gt).norm(dim=-1) * scale).mean()

    @torch.no_grad()
    def log_scalars(self, loss, var, metric=None, step=0, split="train"):
        """model/planar.py:226-254."""
        for key, value in loss.items():
            if key != "all" and self.opt.loss_weight[key] is not None:
                self.tb.add_scalar(f"{split}/loss_{key}", value, step)
        for key, value in (metric or {}).items():
            self.tb.add_scalar(f"{split}/{key}", value, step)
        if self.opt.use_implicit_mask and self.images.get("masks") is not None and not _dist_on():
            self.tb.add_scalar(f"{split}/Mask_Error", self.graph.mse_loss(var.mask_prediction_map, self.images.masks), step)
        if self.opt.use_homographies and self.images.get("gt_hom") is not None:
            self.tb.add_scalar(f"{split}/Homography_Error",
                               self.homography_error(self.graph.warp_param.weight, self.images.gt_hom), step)
        self.tb.add_scalar(f"{split}/PSNR", -10 * loss.rgb.log10(), step)

    @torch.no_grad()
    def visualize(self, var, step=0, split="train"):  # pylint: disable=unused-argument
        """model/planar.py:257-292 — dump the full-canvas frame; TensorBoard images when enabled."""
        if not self._is_main():
            return
        import PIL.Image
        frame = self.predict_entire_image()
        PIL.Image.fromarray((frame * 255).byte().permute(1, 2, 0).numpy()).save(f"{self.vis_path}/{self.vis_it}.png")
        self.vis_it += 1
        if self.tb:
            if self.vis_it == 1:
                self.tb.add_images("train/input_images", var.images.rgb.cpu().clamp(0, 1), self.it + 1)
                if self.opt.use_masks and var.images.get("masks") is not None:
                    self.tb.add_images("train/input_masks", var.images.masks.cpu(), self.it + 1)
            self.tb.add_image("train/predicted_image", frame.clamp(0, 1), self.it + 1)
            if self.opt.use_implicit_mask:
                self.tb.add_images("train/implicit_masks", var.mask_prediction_map.cpu().clamp(0, 1), self.it + 1)
