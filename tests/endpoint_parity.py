"""End-point parity of the training loop (SURVEY.md section 8 row g / BASELINE.json north_star):
patch-corner alignment error (px) and PSNR after `max_iter` iterations of model/planar.py:136-209, this library
against the oracle port of the reference, same synthetic scene, same seed-matched initialisation, same GPU.

Arms (each from the same init, seed 3, Adam 1e-3 x3, fix_first):
  oracle_fp32   oracle/planar_oracle.py (bit-faithful to the reference, pinned by tests/golden) in eager PyTorch fp32 on cuda:0
  oracle_fp64   the same op sequence in float64: the yardstick for how far two faithful fp32 runs may drift apart
                (3000 Adam steps of a ReLU network amplify last-bit differences; |oracle_fp32 - oracle_fp64| is the
                noise floor any other arithmetic order - cuBLAS vs our kernels - has to be read against)
  repo_fp32     marf_b200 precision=fp32
  repo_bf16     marf_b200 precision=bf16
Scenes: `c1` (config-1-like: barf_c2f=[0,0.4], occluders + disk masks), `c2` (config 2: full posenc, learned mask +
edge term), `c3` (config 3: config 2 + occluders pasted in; reports Mask_Error, model/planar.py:237-242).

TEST INFRASTRUCTURE: this file executes the oracle, so it lives under tests/ (tests/test_gpu_endpoint.py runs a short
version; profiles/tools/endpoint_parity.py is a launcher for the full 3000-iteration run kept under profiles/).
    python tests/endpoint_parity.py --scenes c1,c2,c3 --iters 3000 --out profiles/r02_endpoint_parity.json
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

SCENES = {
    # name: (barf_c2f, occluders, implicit mask, edges, disk masks)
    "c1": dict(c2f=[0.0, 0.4], occluders=True, implicit=False, edges=True, masks=True),
    "c2": dict(c2f=None, occluders=False, implicit=True, edges=True, masks=True),
    "c3": dict(c2f=None, occluders=True, implicit=True, edges=True, masks=True),
    # BARF-style run without any mask or edge term (the r01 convergence scene)
    "c0": dict(c2f=[0.0, 0.4], occluders=False, implicit=False, edges=False, masks=False),
}


def make_opt(scene, iters, H, W, precision, args):
    from marf_b200 import options
    s = SCENES[scene]
    opt = options.load_options(os.path.join(ROOT, "options/planar.yaml"))
    opt.update(model="planar", yaml="planar", H=H, W=W, patch_H=H // 2, patch_W=W // 2, batch_size=args.batch,
               use_masks=s["masks"], use_implicit_mask=s["implicit"], use_edges=s["edges"], barf_c2f=s["c2f"],
               max_iter=iters, use_homographies=False, precision=precision, device="cuda:0",
               output_path=f"/tmp/marf_endpoint_{scene}_{precision}", tb=None, seed=3, world_size=1, rank=0,
               fused_optimizer=not args.torch_adam)
    opt.warp.noise_h, opt.warp.noise_t = args.noise_h, args.noise_t
    opt.synthetic = dict(enabled=True, seed=args.scene_seed, occluders=s["occluders"])
    opt.freq.scalar = 10 ** 9
    opt.freq.vis = 10 ** 9
    return opt


def corner_px(cfg_or_opt, warp, gt_warp, H, W):
    import planar_oracle as po
    cfg = po.PlanarConfig(H=H, W=W, patch_H=H // 2, patch_W=W // 2, batch_size=warp.shape[0])
    with torch.device(warp.device):
        pred = po.warp_corners(cfg, warp.detach().float())
        gt = po.warp_corners(cfg, gt_warp.to(warp.device).float())
    return float(((pred - gt).norm(dim=-1) * (max(H, W) / 2.0)).mean())


def summarise(tail_rgb, hist, extra):
    psnr_tail = float(-10 * torch.stack(tail_rgb).double().mean().log10()) if tail_rgb else float("nan")
    out = dict(corner_px=hist[-1]["corner_px"], psnr_final=hist[-1]["psnr"], psnr_last200=psnr_tail, hist=hist)
    out.update(extra)
    return out


def run_repo(scene, iters, H, W, precision, args, every):
    from marf_b200 import planar
    from marf_b200.attrdict import AttrDict
    opt = make_opt(scene, iters, H, W, precision, args)
    os.makedirs(opt.output_path, exist_ok=True)
    torch.manual_seed(3)
    m = planar.Model(opt)
    m.load_dataset()
    m.build_networks()
    m.setup_optimizer()
    m.vis_path = opt.output_path
    m.timer = AttrDict(start=time.time(), it_mean=None)
    var = AttrDict(idx=torch.arange(opt.batch_size), images=m.images)
    hist, tail = [], []
    n_tail = min(200, max(1, iters // 10))
    torch.cuda.synchronize()
    t0 = time.time()
    for it in range(iters):
        loss = m.train_iteration(var, None)
        if it >= iters - n_tail:
            tail.append(loss.rgb.detach().clone())
        if opt.warp.fix_first:
            m.graph.warp_param.weight.data[0] = 0
        if (it + 1) % every == 0 or it == 0 or it == iters - 1:
            hist.append(dict(it=it + 1, corner_px=corner_px(None, m.graph.warp_param.weight.data, m.images.gt_warp, H, W),
                             psnr=float(-10 * loss.rgb.log10()), loss=float(loss.all)))
    torch.cuda.synchronize()
    extra = dict(seconds=time.time() - t0, optimizer="torch.optim.Adam" if args.torch_adam else "marf_adam_step")
    if opt.use_implicit_mask and m.images.get("masks") is not None:
        # Mask_Error (model/planar.py:237-242): mse_loss(mask_prediction_map, images.masks) without a mask argument
        extra["mask_error"] = float(m.graph.mse_loss(var.mask_prediction_map, m.images.masks))
    return summarise(tail, hist, extra), m.images


def run_oracle(scene, iters, H, W, dtype, args, every, images):
    import planar_oracle as po
    s = SCENES[scene]
    dev = "cuda:0"
    cfg = po.PlanarConfig(H=H, W=W, patch_H=H // 2, patch_W=W // 2, batch_size=args.batch, L_2D=8,
                          barf_c2f=tuple(s["c2f"]) if s["c2f"] else None, use_masks=s["masks"],
                          use_implicit_mask=s["implicit"], use_edges=s["edges"], max_iter=iters)
    if s["implicit"] and (cfg.patch_H, cfg.patch_W) != (180, 240):
        pass    # (the oracle's mask path takes any patch size; only the reference hard-wires 180x240, model/planar.py:344)
    p = po.init_params(cfg, seed=3)

    def mv(t):
        return t.to(device=dev, dtype=dtype)
    p.mlp_w, p.mlp_b = [mv(t) for t in p.mlp_w], [mv(t) for t in p.mlp_b]
    p.warp = mv(p.warp)
    if p.mask_w is not None:
        p.mask_w, p.mask_b, p.embed = [mv(t) for t in p.mask_w], [mv(t) for t in p.mask_b], mv(p.embed)
    im = dict(rgb=mv(images.rgb),
              masks=mv(images.masks) if (s["masks"] and images.get("masks") is not None) else None,
              masks_eroded=mv(images.masks_eroded) if (s["masks"] and images.get("masks_eroded") is not None) else None,
              edges=images.edges.to(dev) if (s["edges"] and images.get("edges") is not None) else None)
    groups = [dict(params=list(p.mlp_w) + list(p.mlp_b), lr=1e-3), dict(params=[p.warp], lr=1e-3)]
    if s["implicit"]:
        groups.append(dict(params=list(p.mask_w) + list(p.mask_b), lr=1e-3))
    for g in groups:
        for t in g["params"]:
            t.requires_grad_(True)
    optim = torch.optim.Adam(groups)
    hist, tail = [], []
    n_tail = min(200, max(1, iters // 10))
    progress = 0.0
    out = None
    torch.cuda.synchronize()
    t0 = time.time()
    with torch.device(dev):
        for it in range(iters):
            optim.zero_grad()
            out = po.forward(p, im["rgb"], cfg, progress)
            loss = po.losses(out, im, cfg, it)
            loss["all"].backward()
            optim.step()
            p.warp.data[0] = 0
            progress = (it + 1) / cfg.max_iter
            if it >= iters - n_tail:
                tail.append(loss["rgb"].detach().clone())
            if (it + 1) % every == 0 or it == 0 or it == iters - 1:
                hist.append(dict(it=it + 1, corner_px=corner_px(None, p.warp.data, images.gt_warp, H, W),
                                 psnr=float(-10 * loss["rgb"].detach().log10()), loss=float(loss["all"])))
    torch.cuda.synchronize()
    extra = dict(seconds=time.time() - t0, optimizer="torch.optim.Adam")
    if s["implicit"] and im["masks"] is not None:
        extra["mask_error"] = float(po.mse_loss(out["mask_prediction_map"].detach(), im["masks"]))
    return summarise(tail, hist, extra)


def run_scene(scene, iters, H, W, arms, args, every=250):
    res = {}
    images = None
    for arm in [a for a in arms if a.startswith("repo_")]:
        res[arm], images = run_repo(scene, iters, H, W, arm.split("_")[1], args, every)
        print(f"[{scene}] {arm:12s} corner {res[arm]['corner_px']:.4f} px  PSNR(last200) {res[arm]['psnr_last200']:.2f} dB  "
              f"PSNR(final) {res[arm]['psnr_final']:.2f}  mask_error {res[arm].get('mask_error')}  {res[arm]['seconds']:.1f}s", flush=True)
    if images is None:                     # oracle-only run: build the scene through the same generator
        from marf_b200 import synth
        opt = make_opt(scene, iters, H, W, "fp32", args)
        images = synth.make_scene(opt, seed=args.scene_seed, occluders=SCENES[scene]["occluders"])
    for arm in [a for a in arms if a.startswith("oracle_")]:
        dtype = torch.float64 if arm.endswith("fp64") else torch.float32
        res[arm] = run_oracle(scene, iters, H, W, dtype, args, every, images)
        print(f"[{scene}] {arm:12s} corner {res[arm]['corner_px']:.4f} px  PSNR(last200) {res[arm]['psnr_last200']:.2f} dB  "
              f"PSNR(final) {res[arm]['psnr_final']:.2f}  mask_error {res[arm].get('mask_error')}  {res[arm]['seconds']:.1f}s", flush=True)
    ref = res.get("oracle_fp32")
    if ref:
        for arm, r in res.items():
            if arm != "oracle_fp32":
                r["delta_vs_oracle_fp32"] = dict(corner_px=abs(r["corner_px"] - ref["corner_px"]),
                                                 psnr_last200=abs(r["psnr_last200"] - ref["psnr_last200"]))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", default="c1,c2,c3")
    ap.add_argument("--arms", default="repo_fp32,repo_bf16,oracle_fp32,oracle_fp64")
    ap.add_argument("--iters", type=int, default=3000)
    ap.add_argument("--H", type=int, default=360)
    ap.add_argument("--W", type=int, default=480)
    ap.add_argument("--batch", type=int, default=5)
    ap.add_argument("--noise_h", type=float, default=0.1)
    ap.add_argument("--noise_t", type=float, default=0.2)
    ap.add_argument("--scene_seed", type=int, default=0)
    ap.add_argument("--torch_adam", action="store_true", help="repo arms step with torch.optim.Adam instead of marf_adam_step")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    os.chdir(ROOT)
    torch.backends.cuda.matmul.allow_tf32 = False        # the reference runs torch defaults: fp32 SGEMM, no TF32
    torch.backends.cudnn.allow_tf32 = False
    out = dict(iters=a.iters, H=a.H, W=a.W, batch=a.batch, noise_h=a.noise_h, noise_t=a.noise_t, scene_seed=a.scene_seed,
               gpu=torch.cuda.get_device_name(0), torch=torch.__version__, scenes={})
    for sc in a.scenes.split(","):
        out["scenes"][sc] = run_scene(sc, a.iters, a.H, a.W, a.arms.split(","), a)
        if a.out:
            json.dump(out, open(a.out, "w"), indent=1)
    return out


if __name__ == "__main__":
    main()
