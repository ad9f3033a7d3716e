"""End-to-end convergence check on a synthetic planar scene with known ground-truth warps:
train the drop-in plugin (Model.train_iteration) in fp32 and bf16 mode from the same seed and report the
patch-corner alignment error (px, after removing the gauge fixed by patch 0) and PSNR, the two quantities of
BASELINE.json's parity criterion.   python profiles/tools/convergence.py [--iters 3000] [--modes fp32,bf16]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.chdir(ROOT)


def run(mode, iters, args):
    from marf_b200 import options, planar
    from marf_b200.attrdict import AttrDict
    opt = options.load_options("options/planar.yaml")
    opt.update(model="planar", yaml="planar", H=args.H, W=args.W, patch_H=args.H // 2, patch_W=args.W // 2, batch_size=5,
               use_masks=args.occluders or args.implicit, use_implicit_mask=args.implicit, use_edges=args.implicit,
               barf_c2f=[0.0, 0.4], max_iter=iters,
               use_homographies=False, precision=mode, device="cuda:0", output_path=f"/tmp/marf_conv_{mode}", tb=None, seed=3,
               world_size=1, rank=0, fused_optimizer=True)
    opt.warp.noise_h, opt.warp.noise_t = args.noise_h, args.noise_t
    opt.synthetic = dict(enabled=True, seed=args.scene_seed, occluders=args.occluders or args.implicit)
    opt.freq.scalar = 10 ** 9
    opt.freq.vis = 10 ** 9
    os.makedirs(opt.output_path, exist_ok=True)
    torch.manual_seed(3)
    m = planar.Model(opt)
    m.load_dataset()
    m.build_networks()
    m.setup_optimizer()
    m.vis_path = opt.output_path
    m.timer = AttrDict(start=time.time(), it_mean=None)
    var = AttrDict(idx=torch.arange(opt.batch_size), images=m.images)
    hist = []
    tail = []                                  # rgb loss of the last 200 iterations (PSNR of their mean: a steadier end point)
    t0 = time.time()
    for it in range(iters):
        loss = m.train_iteration(var, None)
        if it >= iters - 200:
            tail.append(loss.rgb.detach())
        if opt.warp.fix_first:
            m.graph.warp_param.weight.data[0] = 0
        if (it + 1) % args.every == 0 or it == 0:
            err = float(m.corner_error_px(m.images.gt_warp))
            psnr = float(-10 * loss.rgb.log10())
            hist.append(dict(it=it + 1, corner_px=err, psnr=psnr, loss=float(loss.all)))
            print(f"[{mode}] it {it+1:5d}  corner error {err:7.3f} px   PSNR {psnr:6.2f} dB   loss {float(loss.all):.5f}", flush=True)
    torch.cuda.synchronize()
    psnr_tail = float(-10 * torch.stack(tail).double().mean().log10()) if tail else float("nan")
    print(f"[{mode}] PSNR of the mean rgb loss over the last {len(tail)} iterations: {psnr_tail:.2f} dB", flush=True)
    return dict(mode=mode, seconds=time.time() - t0, hist=hist, psnr_last200=psnr_tail)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3000)
    ap.add_argument("--every", type=int, default=500)
    ap.add_argument("--modes", default="fp32,bf16")
    ap.add_argument("--H", type=int, default=360)
    ap.add_argument("--W", type=int, default=480)
    ap.add_argument("--noise_h", type=float, default=0.1)
    ap.add_argument("--noise_t", type=float, default=0.2)
    ap.add_argument("--scene_seed", type=int, default=0)
    ap.add_argument("--occluders", action="store_true")
    ap.add_argument("--implicit", action="store_true", help="learned occlusion mask (mask head + edge term), occluders pasted in")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    res = [run(mode, a.iters, a) for mode in a.modes.split(",")]
    if len(res) == 2:
        f, b = res[0]["hist"][-1], res[1]["hist"][-1]
        print(f"final: corner error {res[0]['mode']} {f['corner_px']:.3f} px vs {res[1]['mode']} {b['corner_px']:.3f} px "
              f"(delta {abs(f['corner_px']-b['corner_px']):.3f}); PSNR {f['psnr']:.2f} vs {b['psnr']:.2f} dB "
              f"(delta {abs(f['psnr']-b['psnr']):.2f}); wall {res[0]['seconds']:.1f}s vs {res[1]['seconds']:.1f}s")
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
