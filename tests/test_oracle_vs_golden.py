"""Pin the CPU oracle (oracle/planar_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, made by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import cases
import fixtures as fx
import planar_oracle as po

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))


def test_geometry():
    g = cases.load_golden("unit_geometry")
    cfg = po.PlanarConfig(H=40, W=56, patch_H=20, patch_W=28, batch_size=6)
    h = torch.from_numpy(g["h"])
    assert torch.equal(h, fx.synth_warp(11, 6, scale=0.2, fix_first=True))
    np.testing.assert_allclose(po.sl3_to_SL3(h).numpy(), g["SL3"], rtol=0, atol=1e-7)
    gc = po.normalized_pixel_grid(cfg, crop=True)
    gf = po.normalized_pixel_grid(cfg, crop=False)
    assert np.array_equal(gc.numpy(), g["grid_crop"])          # bit-exact: same f32 op order
    assert np.array_equal(gf.numpy(), g["grid_full"])
    np.testing.assert_allclose(po.warp_grid(gc.repeat(6, 1, 1), h).numpy(), g["warped"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(po.warp_corners(cfg, h).numpy(), g["corners"], rtol=0, atol=2e-7)
    d = po.normalized_pixel_grid(po.PlanarConfig(), crop=True)
    assert np.array_equal(d[:5].numpy(), g["default_grid_first"])
    assert np.array_equal(d[-5:].numpy(), g["default_grid_last"])
    # default x∈[-.4979,.4979], y∈[-.3729,.3729]
    np.testing.assert_allclose(g["default_grid_minmax"], [-0.4979, 0.4979, -0.3729, 0.3729], atol=1e-4)


def test_encodings():
    g = cases.load_golden("unit_encodings")
    coord = torch.from_numpy(g["coord"])
    n = 0
    for key in g:
        if not key.startswith("enc_"):
            continue
        _, Ls, c2fs, ps = key.split("_")
        L = int(Ls[1:])
        c2f = None if c2fs == "c2fnone" else tuple(float(x) for x in c2fs[3:].split("-"))
        prog = float(ps[1:])
        enc = po.positional_encoding(coord, L, c2f, prog)
        np.testing.assert_allclose(enc.numpy(), g[key], rtol=0, atol=1e-6, err_msg=key)
        n += 1
    assert n == 3 * (1 + 2 * 7)
    np.testing.assert_allclose(po.pos_embedding(torch.from_numpy(g["pe_in"])).numpy(), g["pe_out"], rtol=0, atol=1e-6)
    assert g["pe_out"].shape[-1] == 42


def test_stencils_match_opencv_goldens():
    g = cases.load_golden("unit_stencils")
    rgb, masks = fx.synth_patches(7, 2, 23, 31, occluders=True)
    np.testing.assert_allclose(po.sobel_gauss_edges(rgb.numpy()), g["edges3"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(po.sobel_gauss_edges(rgb[:, :1].numpy()), g["edges1"], rtol=0, atol=1e-12)
    assert np.array_equal(po.erode5(masks.numpy()), g["eroded"])
    # and the cv2-calling variant the oracle's forward uses
    np.testing.assert_allclose(po.compute_edges_cv2(rgb).numpy(), g["edges3"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", list(cases.STEP_CASES))
def test_step(name):
    cfg, params, images, it, progress, g = cases.build_case(name)
    out, loss, grads = po.step(params, images, cfg, it=it, progress=progress)
    if cfg.use_edges:
        np.testing.assert_allclose(images["edges"].numpy(), g["edges_label"], rtol=0, atol=1e-12) \
            if "edges_label" in g else None
    if "masks_eroded" in g:
        assert np.array_equal(images["masks_eroded"].numpy(), g["masks_eroded"])
    for k in ("render", "rgb", "mask", "edge", "all"):
        np.testing.assert_allclose(float(loss[k]), float(g["loss_" + k]), rtol=2e-6, atol=1e-9, err_msg=k)
    named = {}
    nl = len(params.mlp_w)
    for i in range(nl):
        named[f"gW{i}"] = grads[i]
        named[f"gb{i}"] = grads[nl + i]
    named["gwarp"] = grads[2 * nl]
    if cfg.use_implicit_mask:
        nm = len(params.mask_w)
        for i in range(nm):
            named[f"gMW{i}"] = grads[2 * nl + 1 + i]
            named[f"gMb{i}"] = grads[2 * nl + 1 + nm + i]
        s = int(g["stride"])
        cases.check_close(out["rgb_prediction"][:, ::s], g["rgb_prediction_s"], 2e-6, "rgb")
        cases.check_close(out["mask_prediction"][:, ::s], g["mask_prediction_s"], 2e-6, "mask")
        cases.check_digest(out["rgb_prediction"], g, "rgb_digest", tol=2e-6)
        cases.check_digest(out["mask_prediction"], g, "mask_digest", tol=2e-6)
        xy = po.normalized_pixel_grid(cfg)
        feats = po.mask_features(images["rgb"][0], xy, params.embed)
        cases.check_close(feats[::997], g["mask_feats_s"], 1e-6, "mask_feats")
    else:
        cases.check_close(out["rgb_prediction"], g["rgb_prediction"], 2e-6, "rgb")
        if cfg.use_edges:
            cases.check_close(out["edge_prediction"], g["edge_prediction"], 1e-5, "edge_pred")
    for k, v in named.items():
        if k in g:
            cases.check_close(v, g[k], 2e-5, k)
        else:
            cases.check_digest(v, g, k + "_digest", tol=2e-5)


@pytest.mark.parametrize("name,over", [
    ("train_small_c2f", dict(cases.SMALL, use_masks=True, barf_c2f=(0.0, 0.4), max_iter=40)),
    ("train_small_edges", dict(cases.SMALL, use_masks=True, use_edges=True, max_iter=40)),
    # the default 4x256 / L=8 network (the shape the bf16 tensor-core path serves): MID size, and the learned mask at full size
    ("train_mid256_c2f", dict(cases.MID, use_masks=True, barf_c2f=(0.0, 0.4), max_iter=40)),
    ("train_mid256_implicit", dict(batch_size=2, use_masks=True, use_implicit_mask=True, use_edges=True, max_iter=40)),
])
def test_training_trajectory(name, over):
    """Model.train_iteration + loop tail (model/planar.py:154-158,187-209): seed-matched init, Adam,
    fix_first, progress schedule, alpha schedule."""
    g = cases.load_golden(name)
    cfg = po.PlanarConfig(**over)
    params = po.init_params(cfg, seed=3)
    np.testing.assert_array_equal(params.mlp_w[0].numpy(), g["init_w0"])
    np.testing.assert_array_equal(params.mlp_w[-1].numpy(), g["init_wl"])
    params.warp = fx.synth_warp(33, cfg.batch_size, scale=0.03)
    images = cases.make_images(cfg, seed=43)
    hist = po.adam_train(params, images, cfg, n_iter=cfg.max_iter)
    for k in ("render", "rgb", "mask", "edge", "all"):
        # (256-wide: 40 Adam steps amplify the fp32 summation-order differences between this run's BLAS threading and the golden run's)
        np.testing.assert_allclose([h[k] for h in hist], g["hist_" + k], rtol=2e-4 if "small" in name else 1e-3,
                                   atol=1e-8 if "small" in name else 1e-5, err_msg=k)
    # (Adam moves a parameter by ~lr = 1e-3 per step whatever the gradient's size: where a warp-gradient entry is near zero its
    #  rounding decides the direction of a step, so the full-posenc 256-wide run pins the warps to about one step)
    np.testing.assert_allclose(params.warp.detach().numpy(), g["warp_final"], rtol=0,
                               atol=2e-5 if "small" in name else (1e-4 if "c2f" in name else 1.5e-3))
    assert float(np.abs(g["warp_final"][0]).max()) == 0.0
