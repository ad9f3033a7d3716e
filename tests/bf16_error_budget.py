"""Where does the bf16 mode's warp-gradient error come from?  CPU emulation of the bf16 path's rounding points
(csrc/bf16_path.cu: activations, weights, dlogits and dY rounded to bf16, fp32 accumulation) on the named test cases,
with individual rounding sources switched off, against the fp32 oracle.  Diagnostic, CPU only:
    python tests/bf16_error_budget.py [case ...]
TEST INFRASTRUCTURE (imports oracle/); not collected by pytest."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import cases  # noqa: E402
import planar_oracle as po  # noqa: E402


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def emulate(cfg, params, images, it, progress, exact=()):
    """exact: subset of {'fwd_act','x0','fwd_w','fwd_w0','dl','dy','dx_w','w0','wlast'} — rounding sources switched OFF."""
    ex = set(exact)
    r_act = (lambda t: t) if "fwd_act" in ex else bf
    r_x0 = (lambda t: t) if ("x0" in ex or "fwd_act" in ex) else bf
    r_fw = (lambda t: t) if "fwd_w" in ex else bf
    r_fw0 = (lambda t: t) if ("fwd_w0" in ex or "fwd_w" in ex) else bf
    r_dl = (lambda t: t) if "dl" in ex else bf
    r_dy = (lambda t: t) if "dy" in ex else bf
    r_dw = (lambda t: t) if "dx_w" in ex else bf
    r_w0 = (lambda t: t) if ("w0" in ex or "dx_w" in ex) else bf
    r_wl = (lambda t: t) if ("wlast" in ex or "dx_w" in ex) else bf
    B, h, w = cfg.batch_size, cfg.h, cfg.w
    warp = params.warp.detach().clone().requires_grad_(True)
    xy = po.normalized_pixel_grid(cfg, crop=cfg.use_cropped_images).repeat(B, 1, 1)
    uv = po.warp_grid(xy, warp)
    enc = po.positional_encoding(uv, cfg.L_2D, cfg.barf_c2f, progress)
    X0f = torch.cat([uv, enc], dim=-1).reshape(B * h * w, -1)
    Ws = [t.detach() for t in params.mlp_w]
    bs = [t.detach() for t in params.mlp_b]
    n = len(Ws)
    X = [r_x0(X0f.detach())]
    for l in range(n - 1):
        z = X[l] @ (r_fw0 if l == 0 else r_fw)(Ws[l]).t() + bs[l]
        X.append(r_act(torch.relu(z)))
    logits = X[n - 1] @ Ws[n - 1].t() + bs[n - 1]          # (output layer: hi/lo bf16 rows of W_last ~ fp32 weights)
    p = torch.sigmoid(logits).reshape(B, h * w, 3)
    # loss gradient wrt logits from the fp32 oracle's own formulas (autograd on the tail only)
    pl = p.detach().clone().requires_grad_(True)
    out = dict(rgb_prediction=pl, rgb_prediction_map=pl.view(B, h, w, 3).permute(0, 3, 1, 2))
    if cfg.use_edges:
        out["edge_prediction"] = torch.from_numpy(po.sobel_gauss_edges(out["rgb_prediction_map"].detach().numpy()))
    loss = po.losses(out, images, cfg, it)
    loss["all"].backward()
    dl = (pl.grad.reshape(-1, 3) * (p.reshape(-1, 3) * (1 - p.reshape(-1, 3)))).float().detach()
    dY = r_dy((r_dl(dl) @ r_wl(Ws[n - 1])) * (X[n - 1] > 0))
    for l in range(n - 2, 0, -1):
        dY = r_dy((dY @ r_dw(Ws[l])) * (X[l] > 0))
    dX0 = dY @ r_w0(Ws[0])
    X0f.backward(dX0)
    return warp.grad.detach(), p.detach()


def main():
    names = sys.argv[1:] or ["mid_mask", "mid_mask_c2f", "mid_nomask_edges"]
    variants = [("all rounding (the shipped path)", ()), ("W0 exact in dX0", ("w0",)), ("dlogits exact", ("dl",)),
                ("W_last exact in dX", ("wlast",)), ("W0 + dl + W_last exact", ("w0", "dl", "wlast")),
                ("all dX weights exact", ("dx_w",)), ("dY exact", ("dy",)), ("dX weights + dY + dl exact (backward fp32)", ("dx_w", "dy", "dl")),
                ("forward exact, backward bf16", ("fwd_act", "fwd_w")),
                ("encoded input X0 exact", ("x0",)), ("forward W0 exact", ("fwd_w0",)), ("X0 + forward W0 exact (layer 0 in fp32)", ("x0", "fwd_w0")),
                ("forward weights exact", ("fwd_w",)), ("forward activations exact", ("fwd_act",)), ("everything exact", ("fwd_act", "fwd_w", "dl", "dy", "dx_w"))]
    for name in names:
        cfg, params, images, it, progress, _ = cases.build_case(name)
        if cfg.use_implicit_mask:
            print(name, ": implicit-mask cases are not emulated here")
            continue
        _, _, grads = po.step(params, images, cfg, it=it, progress=progress)
        gw = grads[2 * len(params.mlp_w)].double()
        print(f"== {name}: |gwarp| = {gw.norm():.3e}")
        for label, ex in variants:
            g, _ = emulate(cfg, params, images, it, progress, ex)
            rel = ((g.double() - gw).norm() / gw.norm()).item()
            print(f"   {label:45s} gwarp rel-L2 {rel:.3e}")


if __name__ == "__main__":
    main()
