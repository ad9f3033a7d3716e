"""The oracle port (the reference's own eager PyTorch op sequence, fp32 + autograd, OpenCV edge round trip) on the SAME B200:
the number a GPU user of the reference gets today.  Informational; bench.py's arms do not use it."""
import os, sys, time, torch
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
os.chdir(ROOT)
import bench, fixtures as fx, planar_oracle as po
wl = bench.workload("config2", 1)
cfg = po.PlanarConfig(H=wl["H"], W=wl["W"], patch_H=wl["patch_H"], patch_W=wl["patch_W"], batch_size=5,
                      layers=tuple([None] + wl["layers"]), L_2D=wl["L"], barf_c2f=wl["c2f"], use_masks=wl["masks"],
                      use_implicit_mask=wl["implicit"], use_edges=wl["edges"])
params = po.init_params(cfg, seed=3)
rgb, masks = fx.synth_patches(0, 5, cfg.h, cfg.w, occluders=True)
images = dict(rgb=rgb, masks=masks, masks_eroded=torch.from_numpy(po.erode5(masks.numpy())), edges=None)
gray = (0.299 * rgb[:, 0:1] + 0.587 * rgb[:, 1:2] + 0.114 * rgb[:, 2:3])
images["edges"] = torch.from_numpy(po.sobel_gauss_edges(gray.numpy()))
dev = "cuda:0"
for k in ("mlp_w", "mlp_b", "mask_w", "mask_b"):
    setattr(params, k, [t.to(dev) for t in getattr(params, k)])
params.warp = params.warp.to(dev); params.embed = params.embed.to(dev)
images = {k: (v.to(dev) if v is not None else None) for k, v in images.items()}
try:
    for i in range(8):
        if i == 3: torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.device(dev):
            po.step(params, images, cfg, it=i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("eager oracle on cuda:0:", dt * 1e3, "ms/step", 216000 / dt / 1e6, "M px/s")
except Exception as e:
    import traceback; traceback.print_exc()
