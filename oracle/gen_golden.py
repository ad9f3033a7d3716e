"""Generate tests/golden/*.npz by running the UNMODIFIED reference (CPU) on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container:  `python oracle/gen_golden.py`
Inputs are rebuilt in the tests from `oracle/fixtures.py` (same seeds), so the fixtures hold
reference OUTPUTS (plus the small inputs where rebuilding would be awkward).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fixtures as fx            # noqa: E402
import ref_runner                # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)

planar, warp_mod, inputs_mod, edict = ref_runner.load_reference()


def save(name, **arrs):
    flat = {}
    for k, v in arrs.items():
        if isinstance(v, dict):       # digest
            for kk, vv in v.items():
                flat[f"{k}.{kk}"] = np.asarray(vv)
        elif torch.is_tensor(v):
            flat[k] = v.detach().cpu().numpy()
        else:
            flat[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"wrote {path}  ({os.path.getsize(path)/1024:.1f} KiB)")


# ------------------------------------------------------------------ unit: geometry
def unit_geometry():
    opt = ref_runner.make_opt(edict, H=40, W=56, patch_H=20, patch_W=28, batch_size=6)
    W = warp_mod.Warp(opt)
    h = fx.synth_warp(11, 6, scale=0.2, fix_first=True)
    H = warp_mod.lie.sl3_to_SL3(h)
    grid_c = W.get_normalized_pixel_grid(crop=True)
    grid_f = W.get_normalized_pixel_grid(crop=False)
    warped = W.warp_grid(grid_c, h)
    corners = W.warp_corners(h)
    opt2 = ref_runner.make_opt(edict)   # planar.yaml defaults
    W2 = warp_mod.Warp(opt2)
    g2 = W2.get_normalized_pixel_grid(crop=True)[0]
    save("unit_geometry", h=h, SL3=H, grid_crop=grid_c[0], grid_full=grid_f[0], warped=warped, corners=corners,
         default_grid_first=g2[:5], default_grid_last=g2[-5:],
         default_grid_minmax=np.array([g2[:, 0].min(), g2[:, 0].max(), g2[:, 1].min(), g2[:, 1].max()]))


# ------------------------------------------------------------------ unit: encodings
def unit_encodings():
    rs = np.random.RandomState(5)
    coord = torch.from_numpy(rs.uniform(-0.7, 0.7, size=(3, 40, 2)).astype(np.float32))
    out = {"coord": coord}
    for L in (8, 4, 10):
        for c2f in (None, (0.0, 0.4), (0.1, 0.5)):
            for prog in (0.0, 0.05, 0.17, 0.25, 0.4, 0.77, 1.0):
                if c2f is None and prog != 0.0:
                    continue
                opt = ref_runner.make_opt(edict, barf_c2f=list(c2f) if c2f else None,
                                          arch=dict(posenc=dict(L_2D=L), layers=[None, 16, 3]))
                ni = planar.NeuralImageFunction(opt)
                ni.progress.data.fill_(prog)
                enc = ni.positional_encoding(coord)
                tag = f"L{L}_c2f{'none' if c2f is None else f'{c2f[0]}-{c2f[1]}'}_p{prog}"
                out["enc_" + tag] = enc
    pe = planar.PosEmbedding(10 - 1, 10)
    xy = torch.from_numpy(rs.uniform(-0.5, 0.5, size=(37, 2)).astype(np.float32))
    out["pe_in"] = xy
    out["pe_out"] = pe(xy)
    save("unit_encodings", **out)


# ------------------------------------------------------------------ helpers for whole-graph cases
def build_graph(opt, seed_w, implicit=False, weight_scale=1.0):
    torch.manual_seed(0)
    g = planar.Graph(opt)
    shapes = [tuple(l.weight.shape) for l in g.neural_image.mlp]
    ws, bs = fx.synth_mlp(seed_w, shapes, scale=weight_scale)
    for l, w, b in zip(g.neural_image.mlp, ws, bs):
        l.weight.data.copy_(w)
        l.bias.data.copy_(b)
    if implicit:
        lins = [m for m in g.implicit_mask.mask_mapping if isinstance(m, torch.nn.Linear)]
        mshapes = [tuple(l.weight.shape) for l in lins]
        mws, mbs = fx.synth_mlp(seed_w + 1, mshapes)
        for l, w, b in zip(lins, mws, mbs):
            l.weight.data.copy_(w)
            l.bias.data.copy_(b)
        g.embedding_view.weight.data.copy_(fx.synth_embed(seed_w + 2, opt.N_vocab, 128))
    return g


def make_images(opt, seed):
    B, h, w = opt.batch_size, opt.patch_H, opt.patch_W
    rgb, masks = fx.synth_patches(seed, B, h, w, occluders=True)
    gray = (0.299 * rgb[:, 0:1] + 0.587 * rgb[:, 1:2] + 0.114 * rgb[:, 2:3])
    gray = torch.round(gray * 255) / 255
    im = edict(rgb=rgb, gray=gray)
    im.masks = masks if opt.use_masks else None
    im.masks_eroded = inputs_mod.erode_images(masks, "cpu", kernel=(5, 5)) if opt.use_masks else None
    im.edges = inputs_mod.compute_edges(gray, "cpu") if opt.use_edges else None
    return im


def run_step(g, opt, images, it, progress):
    g.it = it
    g.neural_image.progress.data.fill_(progress)
    g.zero_grad()
    var = edict(idx=torch.arange(opt.batch_size), images=images)
    var = g.forward(var, mode="train")
    loss = g.compute_loss(var, mode="train")
    # Model.summarize_loss (model/planar.py:172-185) needs a Model; call it unbound with a light stand-in
    holder = type("M", (), {"opt": opt})()
    loss = planar.Model.summarize_loss(holder, loss)
    loss.all.backward()
    return var, loss


def grads_of(g, implicit):
    out = {}
    for i, l in enumerate(g.neural_image.mlp):
        out[f"gW{i}"] = l.weight.grad
        out[f"gb{i}"] = l.bias.grad
    out["gwarp"] = g.warp_param.weight.grad
    if implicit:
        lins = [m for m in g.implicit_mask.mask_mapping if isinstance(m, torch.nn.Linear)]
        for i, l in enumerate(lins):
            out[f"gMW{i}"] = l.weight.grad
            out[f"gMb{i}"] = l.bias.grad
    return out


SMALL = dict(H=40, W=56, patch_H=20, patch_W=28, batch_size=3, max_iter=200,
             arch=dict(layers=[None, 64, 64, 64, 3], skip=[], posenc=dict(L_2D=4)))
MID = dict(H=72, W=96, patch_H=36, patch_W=48, batch_size=3, max_iter=3000)   # default 4x256, L=8


def step_cases():
    variants = {
        "small_nomask": dict(SMALL, use_masks=False),
        "small_mask": dict(SMALL, use_masks=True),
        "small_mask_c2f": dict(SMALL, use_masks=True, barf_c2f=[0.0, 0.4]),
        "small_mask_edges": dict(SMALL, use_masks=True, use_edges=True),
        "small_noposenc": dict(SMALL, use_masks=True, arch=dict(layers=[None, 64, 64, 64, 3], skip=[], posenc=None)),
        "small_skip": dict(SMALL, use_masks=True, arch=dict(layers=[None, 64, 64, 64, 3], skip=[2], posenc=dict(L_2D=4))),
        "small_weights": dict(SMALL, use_masks=True, use_edges=True, loss_weight=dict(render=0, rgb=-1, edge=0.5, mask=None)),
        "mid_mask": dict(MID, use_masks=True),
        "mid_mask_c2f": dict(MID, use_masks=True, barf_c2f=[0.0, 0.4]),
        "mid_nomask_edges": dict(MID, use_masks=False, use_edges=True),
    }
    for name, over in variants.items():
        opt = ref_runner.make_opt(edict, **over)
        g = build_graph(opt, seed_w=21, weight_scale=2.0)
        g.warp_param.weight.data.copy_(fx.synth_warp(31, opt.batch_size, scale=0.05))
        images = make_images(opt, seed=41)
        it = 60 if opt.max_iter == 200 else 450
        progress = it / opt.max_iter
        var, loss = run_step(g, opt, images, it=it, progress=progress)
        arrs = dict(it=it, progress=progress, rgb_prediction=var.rgb_prediction,
                    edge_prediction=var.edge_prediction if opt.use_edges else np.zeros(1),
                    **{f"loss_{k}": np.float64(float(v)) for k, v in loss.items()})
        for k, v in grads_of(g, implicit=False).items():
            if v.numel() <= 20000:
                arrs[k] = v
            else:
                arrs[k + "_digest"] = fx.digest(v)
        if opt.use_edges:
            arrs["edges_label"] = images.edges
        if opt.use_masks:
            arrs["masks_eroded"] = images.masks_eroded
        save("step_" + name, **arrs)


def implicit_cases():
    """The reference hard-codes 180x240 in the mask path (model/planar.py:344) -> full-size patches, B=2."""
    for name, over in {
        "implicit": dict(batch_size=2, use_masks=True, use_implicit_mask=True),
        "implicit_edges": dict(batch_size=2, use_masks=True, use_implicit_mask=True, use_edges=True),
    }.items():
        opt = ref_runner.make_opt(edict, **over)
        g = build_graph(opt, seed_w=22, implicit=True, weight_scale=2.0)
        g.warp_param.weight.data.copy_(fx.synth_warp(32, opt.batch_size, scale=0.05))
        images = make_images(opt, seed=42)
        it = 450
        t0 = time.time()
        var, loss = run_step(g, opt, images, it=it, progress=it / opt.max_iter)
        print(f"  {name}: reference step took {time.time()-t0:.1f}s")
        stride = 37
        arrs = dict(it=it, progress=it / opt.max_iter, stride=stride,
                    rgb_prediction_s=var.rgb_prediction[:, ::stride],
                    mask_prediction_s=var.mask_prediction[:, ::stride],
                    rgb_digest=fx.digest(var.rgb_prediction), mask_digest=fx.digest(var.mask_prediction),
                    **{f"loss_{k}": np.float64(float(v)) for k, v in loss.items()})
        for k, v in grads_of(g, implicit=True).items():
            if v.numel() <= 4096:
                arrs[k] = v
            else:
                arrs[k + "_digest"] = fx.digest(v)
        # the mask head's constant input features for patch 0 (model/planar.py:342-349), subsampled
        xy = g.warp.get_normalized_pixel_grid(crop=True)
        flat = images.rgb[0].long().view(3, -1).permute(1, 0)
        feats = torch.cat((g.embedding_view(flat).view(180, 240, 3, -1).view(-1, 384), g.embedding_uv(xy[0])), dim=-1)
        arrs["mask_feats_s"] = feats[::997]
        save("step_" + name, **arrs)


class _Loader:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def set_postfix(self, **kw):
        pass


def train_cases(only=None):
    """Drive the reference's own Model.train_iteration + the loop tail (model/planar.py:154-158)."""
    for name, over in {
        "train_small_c2f": dict(SMALL, use_masks=True, barf_c2f=[0.0, 0.4], max_iter=40),
        "train_small_edges": dict(SMALL, use_masks=True, use_edges=True, max_iter=40),
        # the default 4x256 / L=8 network at the MID size: the shape the bf16 tensor-core path serves (bf16 trajectory pin)
        "train_mid256_c2f": dict(MID, use_masks=True, barf_c2f=[0.0, 0.4], max_iter=40),
        "train_mid256_implicit": dict(batch_size=2, use_masks=True, use_implicit_mask=True, use_edges=True, max_iter=40),
    }.items():
        if only and name not in only:
            continue
        opt = ref_runner.make_opt(edict, **over)
        opt.output_path = "/tmp/marf_ref_out/" + name
        torch.manual_seed(3)
        m = planar.Model(opt)
        m.images = make_images(opt, seed=43)
        # the reference builds its own seeded weights here (seed-matched init parity is checked on these)
        m.build_networks()
        init_w0 = m.graph.neural_image.mlp[0].weight.detach().clone()
        init_wl = m.graph.neural_image.mlp[-1].weight.detach().clone()
        # start from a non-trivial misalignment so the warp gradient matters
        m.graph.warp_param.weight.data.copy_(fx.synth_warp(33, opt.batch_size, scale=0.03))
        m.setup_optimizer()
        m.timer = edict(start=time.time(), it_mean=None)
        m.graph.train()
        var = edict(idx=torch.arange(opt.batch_size), images=m.images)
        loader = _Loader(opt.max_iter)
        hist = {k: [] for k in ("render", "rgb", "mask", "edge", "all")}
        warps = []
        for _ in range(opt.max_iter):
            loss = m.train_iteration(var, loader)
            if opt.warp.fix_first:
                m.graph.warp_param.weight.data[0] = 0
            for k in hist:
                hist[k].append(float(loss[k]))
            warps.append(m.graph.warp_param.weight.detach().clone())
        save(name, init_w0=init_w0, init_wl=init_wl, warp_final=warps[-1], warp_it10=warps[9],
             w0_final=m.graph.neural_image.mlp[0].weight, progress_final=m.graph.neural_image.progress,
             **{f"hist_{k}": np.array(v) for k, v in hist.items()})


def stencil_cases():
    """cv2 outputs for the numpy restatement of the edge/erode stencils (inputs.py:50-85)."""
    rgb, masks = fx.synth_patches(7, 2, 23, 31, occluders=True)
    edges3 = inputs_mod.compute_edges(rgb, "cpu")
    gray = rgb[:, :1]
    edges1 = inputs_mod.compute_edges(gray, "cpu")
    er = inputs_mod.erode_images(masks, "cpu", kernel=(5, 5))
    save("unit_stencils", edges3=edges3, edges1=edges1, eroded=er)


if __name__ == "__main__":
    which = sys.argv[1:] or ["geometry", "encodings", "stencils", "steps", "implicit", "train"]
    with torch.enable_grad():
        if "geometry" in which:
            unit_geometry()
        if "encodings" in which:
            unit_encodings()
        if "stencils" in which:
            stencil_cases()
        if "steps" in which:
            step_cases()
        if "implicit" in which:
            implicit_cases()
        if "train" in which:
            train_cases()
        if "train256" in which:
            train_cases(only=("train_mid256_c2f", "train_mid256_implicit"))
