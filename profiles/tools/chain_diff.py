"""Debug helper: where do the fused chain kernel and the per-layer kernels disagree (rows -> tiles -> items)?"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import cases, gpu_util

name = sys.argv[1] if len(sys.argv) > 1 else "implicit"
cfg, params, images, it, progress, g = cases.build_case(name)
def grab(eng):
    import ctypes as C
    rows = (cfg.batch_size * cfg.h * cfg.w + 127) // 128 * 128
    out = {}
    for chain in range(2 if cfg.use_implicit_mask else 1):
        for which in (0, 1):
            for layer in range(1, 5) if which == 0 else range(0, 4):
                t = torch.zeros(rows, 256, device="cuda")
                rc = eng.lib.marf_debug_read_bf16(eng.handle, chain, which, layer, C.c_void_p(t.data_ptr()), rows, None)
                assert rc == 0, rc
                out[(chain, which, layer)] = t.cpu()
    return out

eng = gpu_util.make_engine(cfg, "bf16")
fused = gpu_util.run_step(eng, cfg, params, images, it, progress)
fa = grab(eng)
eng.close()
os.environ["MARF_NO_FUSE"] = "1"
eng = gpu_util.make_engine(cfg, "bf16")
plain = gpu_util.run_step(eng, cfg, params, images, it, progress)
pa = grab(eng)
eng.close()
np.set_printoptions(linewidth=200, precision=3, suppress=True)
for k in sorted(fa):
    d = (fa[k] - pa[k]).abs()
    tile_err = d.reshape(-1, 128, 4, 64).amax(dim=(1, 3))      # [tiles, slabs]
    bad = (tile_err > 1e-3).nonzero()
    print("chain %d %s layer %d: max %.4g, bad (tile,slab) pairs %d" % (k[0], "act" if k[1] == 0 else "dY ", k[2], d.max().item(), len(bad)),
          "first:", bad[:6].tolist())
for key in ("rgb_pred", "mask_pred"):
    if fused[key] is None: continue
    d = (fused[key] - plain[key]).abs().reshape(-1, fused[key].shape[-1]).max(dim=1).values.numpy()
    bad = np.nonzero(d > 1e-3)[0]
    print(key, "max", d.max(), "bad rows", len(bad), "of", len(d))
    if len(bad):
        tiles = np.unique(bad // 128)
        print(" bad tiles:", tiles[:40], "... n =", len(tiles))
        print(" bad pairs %148:", np.unique((tiles // 2) % 148)[:40])
        print(" items (pair) :", np.unique(tiles // 2)[:60])
        t0 = tiles[0]
        np.set_printoptions(linewidth=200, precision=3, suppress=True)
        print(" first bad tile", t0, "per-row err:\n", d[t0 * 128:(t0 + 1) * 128])
        f = fused[key].reshape(-1, fused[key].shape[-1]).numpy(); p_ = plain[key].reshape(-1, plain[key].shape[-1]).numpy()
        print(" fused rows:", f[t0 * 128:t0 * 128 + 4], "\n plain rows:", p_[t0 * 128:t0 * 128 + 4])
for k, v in plain["grads"].items():
    rel = ((fused["grads"][k].double() - v.double()).norm() / (v.double().norm() + 1e-30)).item()
    if rel > 5e-3: print("grad", k, rel)

# hypotheses for the first bad tile of chain 0, layer-1 activation
k = (0, 0, 1)
d = (fa[k] - pa[k]).abs().reshape(-1, 128, 256).amax(dim=(1, 2))
bad = (d > 1e-3).nonzero().flatten()
if len(bad):
    t = int(bad[0])
    F = fa[k].reshape(-1, 128, 256); P = pa[k].reshape(-1, 128, 256)
    print("tile", t, "fused row0[:8]", F[t, 0, :8].numpy(), "\n plain row0[:8]", P[t, 0, :8].numpy())
    print(" fused row0[64:72]", F[t, 0, 64:72].numpy(), "\n plain row0[64:72]", P[t, 0, 64:72].numpy())
    for name, cand in (("same pair tile0", P[t - 1]), ("prev item tile1 (t-296)", P[t - 296]), ("prev item tile0", P[t - 297])):
        print(" vs", name, "max diff on cols<128:", (F[t, :, :128] - cand[:, :128]).abs().max().item())
    for L in (2, 3, 4):
        Pl = pa[(0, 0, L)].reshape(-1, 128, 256)
        for name, tt in (("prev item tile1 act%d" % L, t - 296), ("same tile act%d" % L, t)):
            print(" vs", name, (F[t, :, :128] - Pl[tt][:, :128]).abs().max().item())
    print(" frac zero fused", (F[t, :, :128] == 0).float().mean().item(), "plain", (P[t, :, :128] == 0).float().mean().item())
    print(" per-row err[:16]", (F[t, :16, :128] - P[t, :16, :128]).abs().amax(dim=1).numpy())
    print(" per-col err[:16]", (F[t, :, :16] - P[t, :, :16]).abs().amax(dim=0).numpy())
