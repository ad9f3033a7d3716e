"""A genuinely multi-chunk step at BASELINE config 4's patch size: 4 patches of 1024x1024 of a 2048x2048 canvas = 4,194,304
pixel-samples = 4 chunks at the library's DEFAULT chunk size in bf16 mode (2^20; 8 chunks of 2^19 in fp32 mode), disk masks
(static normaliser: forward + backward chunk by chunk in one sweep), default 4x256 / L=8 network, nn.Linear initialisation.
The oracle runs on the same GPU: its autograd keeps ~50 GB of fp32 activations at this size."""
import numpy as np
import pytest
import torch

import cases
import fixtures as fx
import planar_oracle as po

pytestmark = pytest.mark.gpu


def _oracle_on_gpu(cfg, params, images, it, progress):
    dev = "cuda:0"
    p = po.PlanarParams([t.to(dev) for t in params.mlp_w], [t.to(dev) for t in params.mlp_b], params.warp.to(dev))
    im = {k: (v.to(dev) if v is not None else None) for k, v in images.items()}
    with torch.device(dev):
        out, loss, grads = po.step(p, im, cfg, it=it, progress=progress)
    res = dict(rgb=out["rgb_prediction"].detach().cpu(), loss={k: float(v) for k, v in loss.items()}, grads=[g.detach().cpu() for g in grads])
    del out, loss, grads, p, im
    torch.cuda.empty_cache()
    return res


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_step_four_chunks_of_2p20_matches_oracle(precision):
    import gpu_util
    cfg = po.PlanarConfig(H=2048, W=2048, patch_H=1024, patch_W=1024, batch_size=4, use_masks=True, max_iter=3000)
    params = po.init_params(cfg, seed=3)
    params.warp = fx.synth_warp(31, 4, scale=0.05)
    images = cases.make_images(cfg, 41)
    it, progress = 450, 450 / 3000
    ref = _oracle_on_gpu(cfg, params, images, it, progress)
    eng = gpu_util.make_engine(cfg, precision)
    n_chunks = -(-eng.n_local // (1 << 20 if precision == "bf16" else 1 << 19))
    assert n_chunks == (4 if precision == "bf16" else 8)
    res = gpu_util.run_step(eng, cfg, params, images, it, progress)
    out_tol, loss_tol, grad_l2 = (6e-3, 1e-3, 5e-2) if precision == "bf16" else (2e-5, 2e-5, 2e-3)
    assert (res["rgb_pred"] - ref["rgb"]).abs().max().item() <= out_tol
    for k in ("rgb", "all"):
        assert abs(res["losses"][k] - ref["loss"][k]) <= loss_tol * abs(ref["loss"][k]), (k, res["losses"][k], ref["loss"][k])
    nl = len(params.mlp_w)
    named = {f"gW{i}": ref["grads"][i] for i in range(nl)}
    named.update({f"gb{i}": ref["grads"][nl + i] for i in range(nl)})
    named["gwarp"] = ref["grads"][2 * nl]
    report = {}
    for k, v in named.items():
        report[k] = ((res["grads"][k].double() - v.double()).norm() / (v.double().norm() + 1e-30)).item()
    print(precision, {k: f"{v:.2e}" for k, v in report.items()})
    for k, v in report.items():
        # (the warp gradient is a small residual of per-pixel terms that cancel: bf16 1.5e-1 as in test_gpu_parity_bf16, fp32 2e-2)
        bound = grad_l2 if k != "gwarp" else (1.5e-1 if precision == "bf16" else 2e-2)
        assert v <= bound, (k, report)
    assert res["nonfinite"] == 0.0
    eng.close()


def test_patch_shards_add_up_at_config4_patch_size():
    """Data-parallel arithmetic at BASELINE config 4's patch size on ONE GPU: 4 patches of 1024x1024 as one 4-chunk step ==
    the sum of two 2-patch shards (the rank-0 and rank-1 engines of a 2-rank job run one after the other) when both shards use
    the global loss normaliser — gradients are sums over pixel-samples.  bf16 tensor-core mode; what remains
    is the order of the fp32 reductions."""
    import gpu_util
    from marf_b200 import _lib as L
    cfg = po.PlanarConfig(H=2048, W=2048, patch_H=1024, patch_W=1024, batch_size=4, use_masks=True, max_iter=3000)
    params = po.init_params(cfg, seed=3)
    params.warp = fx.synth_warp(31, 4, scale=0.05)
    images = cases.make_images(cfg, 41)
    it, progress = 450, 450 / 3000
    norm = 3.0 * float(images["masks"].double().sum())

    def run(rank, world):
        eng = gpu_util.make_engine(cfg, "bf16", rank=rank, world=world)
        kw = gpu_util.step_kwargs(eng, cfg, params, images, it, progress)
        lo, n = eng.patch_offset, eng.batch
        for k in ("rgb", "masks"):
            kw[k] = kw[k][lo:lo + n].contiguous()
        kw["rgb_pred"] = torch.zeros(n, cfg.h * cfg.w, 3, device=eng.device)
        kw["norm_rgb"] = norm
        eng.step(**kw)
        torch.cuda.synchronize()
        flat = torch.cat([t.flatten() for t in kw["g_mlp_w"] + kw["g_mlp_b"] + [kw["g_warp"]]]).double().cpu()
        sums = eng.sums.clone().cpu()
        eng.close()
        return flat, sums

    whole, s_whole = run(0, 1)
    a, s_a = run(0, 2)
    b, s_b = run(1, 2)
    rel = ((a + b - whole).norm() / whole.norm()).item()
    print("shards vs whole: gradient rel-L2", rel)
    assert rel <= 2e-5, rel
    assert abs(float(s_a[L.S_RGB] + s_b[L.S_RGB]) - float(s_whole[L.S_RGB])) <= 1e-9 * float(s_whole[L.S_RGB])
    assert float(s_a[L.N_RGB] + s_b[L.N_RGB]) == float(s_whole[L.N_RGB])
