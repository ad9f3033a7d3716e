"""bench.py — pixel-samples/s of MARF's planar bundle-adjusting training step (fwd + bwd + warp grad).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl marf|reference] [--precision fp32|bf16]
                    [--workload config2|config4|config5]

N=1 workload (BASELINE.json configs[1]): planar.yaml, full positional encoding (no barf_c2f), implicit mask
network enabled, edge term on (yaml default), synthetic 360x480 scene, 5 patches of 180x240 -> 216,000
pixel-samples per step.  N>1 (torchrun, one rank per GPU): weak scaling — every rank owns 5 such patches
(B = 5N), gradients + loss sums all-reduced over NCCL every step.

`value`  : whole-job pixel-samples/s of marf_step (+ all-reduce for N>1) with inputs resident in HBM, timed
           with CUDA events per step on the launching stream, L2 flushed between steps, max over ranks.
`e2e`    : the same metric through the reference-facing plugin (Model.train_iteration incl. Adam), with the
           step's targets copied host->device from pinned memory and the loss read back, every step.
`--impl reference` times the CPU oracle port (oracle/planar_oracle.py, eager PyTorch fp32 on all host cores —
the reference itself is Python and cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PX = {  # algorithmic: 2 * MACs of fwd + dX + dW
    "mlp256_L8": 1_236_480, "mlp256_L8_mask": 2_853_888, "mlp512_L10": 4_856_832,
}


def workload(name, n_gpus):
    wl = _workload(name, n_gpus)
    wl["batch_per_gpu"] = wl["batch"] // n_gpus if wl["scaling"] == "strong" else wl["batch"] // n_gpus
    return wl


def _workload(name, n_gpus):
    if name == "config2":
        return dict(name="config2: planar.yaml full posenc + implicit mask network + edge term, synthetic 360x480, "
                         f"{5 * n_gpus} patches 180x240 ({5} per GPU)",
                    H=360, W=480, patch_H=180, patch_W=240, batch=5 * n_gpus, layers=[256, 256, 256, 256, 3], L=8,
                    c2f=None, implicit=True, masks=True, edges=True, flop=FLOP_PER_PX["mlp256_L8_mask"], scaling="weak")
    if name == "config4":
        return dict(name="config4: 64 patches 1024x1024 of a synthetic 2048x2048 image, data-parallel",
                    H=2048, W=2048, patch_H=1024, patch_W=1024, batch=64, layers=[256, 256, 256, 256, 3], L=8,
                    c2f=None, implicit=False, masks=True, edges=False, flop=FLOP_PER_PX["mlp256_L8"], scaling="strong")
    if name == "config5":
        return dict(name="config5: width-512 MLP, L=10, 256 patches 512x512 of a synthetic 4096x4096 image "
                         "(patch reduced from 2048x2048 for run time; FLOP/px unchanged)",
                    H=4096, W=4096, patch_H=512, patch_W=512, batch=256, layers=[512, 512, 512, 512, 3], L=10,
                    c2f=None, implicit=False, masks=True, edges=False, flop=FLOP_PER_PX["mlp512_L10"], scaling="strong")
    raise SystemExit(f"unknown workload {name}")


def make_opt(wl, device, precision, out_dir):
    from marf_b200 import options
    opt = options.load_options("options/planar.yaml")
    opt.update(model="planar", yaml="planar", H=wl["H"], W=wl["W"], patch_H=wl["patch_H"], patch_W=wl["patch_W"],
               batch_size=wl["batch"], use_masks=wl["masks"], use_implicit_mask=wl["implicit"], use_edges=wl["edges"],
               barf_c2f=wl["c2f"], use_homographies=False, precision=precision, device=device, output_path=out_dir,
               tb=None, seed=3, world_size=int(os.environ.get("WORLD_SIZE", "1")), rank=int(os.environ.get("RANK", "0")))
    opt.arch.layers = [None] + wl["layers"]
    opt.arch.posenc.L_2D = wl["L"]
    opt.synthetic = dict(enabled=True, seed=0, occluders=True)
    opt.fused_optimizer = True
    if os.environ.get("MARF_BENCH_CHUNK"):              # (experiments: pixel-samples per pass, library default 2^20 in bf16 mode)
        opt.max_chunk_pixels = int(os.environ["MARF_BENCH_CHUNK"])
    opt.freq.scalar = 10 ** 9
    opt.freq.vis = 10 ** 9
    return opt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")



# ------------------------------------------------------------------------------------------------ roofline
def kernel_tables(wl):
    """Algorithmic FLOP and HBM bytes per pixel-sample row of every tensor-core kernel class (SURVEY.md
    section 8a/8d: 2 * MACs, unpadded; coordinates are analytic so the step's only inputs are the targets / masks).
      k_tc_chain<fwd>  : 64-wide input in (128 B), 4 activations (2048 B) + 4 mask-bit rows (128 B) + logits (16 B) out
      k_tc_bwd         : every X_l once (dW), the dlogits tiles, the mask bits; dY_0 of the image chain out (warp-gradient
                         GEMM); the dY_l hand-over between its chain pairs and dW pairs goes through L2, not counted
      k_tc_chain<dx> / k_tc_dw : the two halves of k_tc_bwd as separate launches (MARF_NO_BWD_FUSE, 512-wide networks)
      k_tc_gemm<64,wg> : reads dY_0 (512 B)"""
    chains = 2 if wl["implicit"] else 1
    width = wl["layers"][0]
    hidden = len(wl["layers"]) - 1
    k0 = 2 + 4 * wl["L"]
    out_w = [3, 1][:chains]
    f_fwd = sum(2 * ((k0 if c == 0 else 42 + 384) * width + (hidden - 1) * width * width + width * out_w[c]) for c in range(chains))
    f_dx = sum(2 * ((hidden - 1) * width * width + width * out_w[c]) for c in range(chains)) + 2 * k0 * width     # + dX0 of the image MLP
    f_dw = f_fwd
    byt = {"k_tc_dw": chains * ((hidden - 1) * 4 * width + (2 * width + 16) + (2 * width + 128)),
           "k_tc_chain<fwd>": chains * (128 + hidden * 2 * width + hidden * width // 8 + 16),
           "k_tc_chain<dx>": chains * (16 + hidden * width // 8 + hidden * 2 * width),
           "k_tc_gemm<64,warp_grad>": 2 * width}
    byt["k_tc_bwd"] = chains * (hidden * 2 * width + 128 + 2 * 16 + hidden * width // 8) + 2 * width
    flop = {"k_tc_chain<fwd>": f_fwd, "k_tc_dw": f_dw, "k_tc_chain<dx>": f_dx - 2 * k0 * width, "k_tc_gemm<64,warp_grad>": 2 * k0 * width}
    flop["k_tc_bwd"] = flop["k_tc_dw"] + flop["k_tc_chain<dx>"]
    return flop, byt


def load_traffic(workload_key, precision):
    """DRAM bytes per launch per kernel from the committed ncu capture (profiles/r02_kernel_traffic.json), when it holds this workload."""
    for name in ("r02_kernel_traffic.json", "r02_kernel_traffic_config2.json", "r01_kernel_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("workload") == workload_key and tj.get("precision") == precision:
                return tj["dram_bytes_per_launch"], name
    return {}, None


def roofline_lines(wl, wl_key, precision, kern, pk, rows_per_gpu, step_tflops, n_gpus):
    """`roofline` (this path is bounded by the tensor cores): the dominant kernel's ALGORITHMIC FLOP per launch /
    its mean launch duration (CUDA events on the launching stream, per-kernel pass) against the measured bf16 peak (burst: the
    stricter denominator), with the algorithmic bytes and the DRAM traffic of the ncu capture beside it so that the waste ratio
    is in the record; `hbm_view` = the same launch against the HBM roofline; `kernels` = every tensor-core kernel class;
    `step_roofline` = algorithmic FLOP of the whole step / step time."""
    out = {}
    burst, sust = pk["bf16_burst"], pk["bf16_sustained"]
    out["step_roofline"] = dict(bound="tensor", achieved=step_tflops, peak=burst * n_gpus, unit="TFLOP/s", frac=step_tflops / (burst * n_gpus),
                                frac_of_sustained=step_tflops / (sust * n_gpus),
                                note="algorithmic FLOP of the whole step / mean step time, all kernels; peak = measured burst bf16")
    if not kern:
        out["roofline"] = dict(out["step_roofline"], traffic=None, peak_source=pk["source"])
        return out
    flop, byt = kernel_tables(wl)
    rows = (rows_per_gpu + 127) // 128 * 128
    traffic, tsrc = load_traffic(wl_key, precision)
    rowsets = []
    for k, v in kern.items():
        us = v["us_per_launch"]
        n_l = max(1.0, v["launches_per_step"])
        alg_bytes = byt[k] * rows / n_l
        gbs = alg_bytes / (us * 1e-6) / 1e9
        tf = flop[k] * rows / n_l / (us * 1e-6) / 1e12
        tr = traffic.get(k)
        rowsets.append(dict(kernel=k, us_per_launch=us, launches_per_step=v["launches_per_step"], bound="tensor",
                            achieved=tf, peak=burst, unit="TFLOP/s", frac=tf / burst, frac_of_sustained=tf / sust,
                            algorithmic_bytes=alg_bytes, traffic=tr, traffic_over_algorithmic=(tr / alg_bytes) if tr else None,
                            hbm_view=dict(achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                                          traffic_gbs=(tr / (us * 1e-6) / 1e9) if tr else None)))
    rowsets.sort(key=lambda r: -r["us_per_launch"] * r["launches_per_step"])
    dom = rowsets[0]
    out["roofline"] = dict(bound="tensor", achieved=dom["achieved"], peak=dom["peak"], unit="TFLOP/s", frac=dom["frac"],
                           frac_of_sustained=dom["frac_of_sustained"], traffic=dom["traffic"], algorithmic_bytes=dom["algorithmic_bytes"],
                           traffic_over_algorithmic=dom["traffic_over_algorithmic"], hbm_view=dom["hbm_view"],
                           kernel=dom["kernel"], us_per_launch=dom["us_per_launch"],
                           peak_source=pk["source"] + ", burst bf16 (cuBLAS 8192^3 best of 10)", traffic_source=tsrc,
                           note="dominant kernel of the step; achieved = algorithmic FLOP per launch / mean launch duration from CUDA "
                                "events on the launching stream (per-kernel pass, L2 flushed between steps)")
    out["kernels"] = rowsets
    return out

# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference(wl, steps, warmup, patches=1):
    """Oracle port (eager PyTorch fp32 on the host cores) on a bounded sample: `patches` of the workload's patches."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fixtures as fx
    import planar_oracle as po
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = po.PlanarConfig(H=wl["H"], W=wl["W"], patch_H=wl["patch_H"], patch_W=wl["patch_W"], batch_size=patches,
                          layers=tuple([None] + wl["layers"]), L_2D=wl["L"], barf_c2f=wl["c2f"], use_masks=wl["masks"],
                          use_implicit_mask=wl["implicit"], use_edges=wl["edges"])
    if wl["implicit"] and (cfg.patch_H, cfg.patch_W) != (180, 240):
        raise SystemExit("the reference's mask path is hard-wired to 180x240 patches")
    params = po.init_params(cfg, seed=3)
    rgb, masks = fx.synth_patches(0, patches, cfg.h, cfg.w, occluders=True)
    images = dict(rgb=rgb, masks=masks if wl["masks"] else None, masks_eroded=None, edges=None)
    if wl["masks"]:
        images["masks_eroded"] = torch.from_numpy(po.erode5(masks.numpy()))
    if wl["edges"]:
        gray = (0.299 * rgb[:, 0:1] + 0.587 * rgb[:, 1:2] + 0.114 * rgb[:, 2:3])
        images["edges"] = torch.from_numpy(po.sobel_gauss_edges(gray.numpy()))
    n_px = patches * cfg.h * cfg.w
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        po.step(params, images, cfg, it=i)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return dict(value=n_px / sec, unit="pixel-samples/s", cores=cores, kind="port", torch_threads=torch.get_num_threads(),
                sample=f"{patches} patch(es) of {cfg.h}x{cfg.w} = {n_px} pixel-samples per step, {len(times)} timed steps, "
                       f"oracle/planar_oracle.py (eager PyTorch {torch.__version__} fp32 + autograd, the reference's own op sequence)",
                ms_per_step=sec * 1e3), n_px


def eager_gpu_reference(wl, device, steps=5, warmup=3):
    """The oracle port (the reference's own eager PyTorch op sequence incl. its per-step OpenCV edge round trip) on the SAME GPU,
    whole workload: the number a GPU user of the reference gets today.  Informational."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fixtures as fx
    import planar_oracle as po
    B = wl["batch"]
    cfg = po.PlanarConfig(H=wl["H"], W=wl["W"], patch_H=wl["patch_H"], patch_W=wl["patch_W"], batch_size=B,
                          layers=tuple([None] + wl["layers"]), L_2D=wl["L"], barf_c2f=wl["c2f"], use_masks=wl["masks"],
                          use_implicit_mask=wl["implicit"], use_edges=wl["edges"])
    params = po.init_params(cfg, seed=3)
    rgb, masks = fx.synth_patches(0, B, cfg.h, cfg.w, occluders=True)
    images = dict(rgb=rgb, masks=masks if wl["masks"] else None, masks_eroded=None, edges=None)
    if wl["masks"]:
        images["masks_eroded"] = torch.from_numpy(po.erode5(masks.numpy()))
    if wl["edges"]:
        gray = (0.299 * rgb[:, 0:1] + 0.587 * rgb[:, 1:2] + 0.114 * rgb[:, 2:3])
        images["edges"] = torch.from_numpy(po.sobel_gauss_edges(gray.numpy()))
    for k in ("mlp_w", "mlp_b", "mask_w", "mask_b"):
        if getattr(params, k) is not None:
            setattr(params, k, [t.to(device) for t in getattr(params, k)])
    params.warp = params.warp.to(device)
    if params.embed is not None:
        params.embed = params.embed.to(device)
    images = {k: (v.to(device) if v is not None else None) for k, v in images.items()}
    t0 = 0.0
    for i in range(warmup + steps):
        if i == warmup:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        with torch.device(device):
            po.step(params, images, cfg, it=i)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    n_px = B * cfg.h * cfg.w
    return dict(value=n_px / sec, unit="pixel-samples/s", ms_per_step=sec * 1e3, kind="port", device=device,
                what=f"oracle/planar_oracle.py (the reference's eager PyTorch op sequence, fp32 + autograd, OpenCV edge round trip per step) "
                     f"on the same GPU, whole workload ({n_px} pixel-samples per step), {steps} timed steps")


def cpu_sample_patches(wl):
    """Patches per step of the CPU arm: the whole workload when it is small (config 2: 5 patches, 216,000 pixel-samples), else a
    bounded sample of about one million pixel-samples (config 4: one 1024x1024 patch; ~4 s per step on the box's host cores)."""
    per = wl["patch_H"] * wl["patch_W"]
    return max(1, min(wl["batch_per_gpu"], (1 << 20) // per))


def run_reference(args):
    """The reference arm: the reference's own op sequence (oracle port, eager PyTorch fp32 + autograd) on the host cores, on the same
    workload / metric / unit.  Rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload, args.gpus)
    patches = cpu_sample_patches(wl)
    warm = min(args.warmup, 3)
    # keep the whole arm within a few minutes: time one step, then size the timed loop
    probe, n_px = cpu_reference(wl, steps=1, warmup=0, patches=patches)
    steps = max(1, min(args.steps, int(150.0 / max(probe["ms_per_step"] * 1e-3, 1e-3)) - warm))
    base, n_px = cpu_reference(wl, steps=steps, warmup=warm, patches=patches)
    line = dict(metric="pixel-samples/sec (fwd+bwd+warp grad)", value=base["value"], unit="pixel-samples/s", impl="reference",
                n_gpus=args.gpus, steps=steps, warmup=warm, ms_per_step=base["ms_per_step"], higher_is_better=True,
                scaling=wl["scaling"], vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=wl["name"], sample=base["sample"]),
                cpu_baseline=dict(value=base["value"], unit="pixel-samples/s", cores=base["cores"], kind="port", sample=base["sample"]),
                e2e=dict(value=base["value"], unit="pixel-samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def measure(wl_key, precision, args, ctx, steps, warmup, e2e=True, kernels=True, sampler=None):
    """One workload through the library: resident-input `value`, `e2e` through Model.train_iteration with host buffers, and the
    per-kernel pass.  Returns a dict; frees the model afterwards."""
    import torch
    import torch.distributed as dist
    from marf_b200 import planar
    from marf_b200.attrdict import AttrDict
    world, rank, device, flush = ctx["world"], ctx["rank"], ctx["device"], ctx["flush"]
    wl = workload(wl_key, args.gpus)
    out_dir = os.path.join("/tmp", f"marf_bench_{os.getpid()}")
    opt = make_opt(wl, device, precision, out_dir)
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(3)
    m = planar.Model(opt)
    m.load_dataset()
    m.build_networks()
    m.setup_optimizer()
    m.vis_path = out_dir
    m.timer = AttrDict(start=time.time(), it_mean=None)
    g = m.graph
    var = AttrDict(idx=torch.arange(opt.batch_size), images=m.images)
    n_px_total = opt.batch_size * wl["patch_H"] * wl["patch_W"]
    st = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident-input arm: marf_step (+ allreduce) only
    for _ in range(max(3, warmup)):
        g.forward(var, mode="train")
    barrier()
    if sampler is not None and rank == 0:                 # (one nvidia-smi poller per job, not one per rank)
        sampler.start()
        time.sleep(0.1)
    launches0 = g.engine.launches
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        flush.zero_()                                   # L2 flush between timed iterations (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        g.forward(var, mode="train")
        e1.record(st)
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = g.engine.launches - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t) / steps
    res = dict(workload=wl["name"], precision=precision, value=n_px_total / (ms_per_step * 1e-3), ms_per_step=ms_per_step, steps=steps,
               pixel_samples_per_step=n_px_total, flop_per_pixel_sample=wl["flop"], gpu_launches=launches, wall_s_timed_loop=wall)
    res["tflops"] = res["value"] * wl["flop"] / 1e12

    # ---------------- e2e arm: plugin call with host buffers (H2D of targets, D2H of the loss) + Adam
    if e2e:
        loc = g._local[1]                                   # this rank's shard of the resident inputs (what the step reads)
        host = {k: v.cpu().pin_memory() for k, v in dict(rgb=loc.rgb).items()}
        h2d = sum(v.numel() * v.element_size() for v in host.values()) * world     # whole job, bytes per step
        d2h = 8 * world
        # Input prefetch, as a training data loader does it: the targets of step i+1 are copied host->device on a copy stream into
        # the second of two device buffers while step i computes (every step's H2D copy is inside the timed region; a buffer is
        # refilled only after the step that read it has finished).  The loss of every step is read back (D2H into pinned
        # memory); the host consumes it two steps later so that Python/launch overhead overlaps the GPU's work on the next steps.
        copy_st = torch.cuda.Stream(device=device)
        bufs = [loc.rgb, torch.empty_like(loc.rgb)]
        ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
        ev_free = [torch.cuda.Event(), torch.cuda.Event()]
        for ev in ev_free:
            ev.record(st)
        LAG = 2                                             # the host reads step i's loss while steps i+1, i+2 are queued
        loss_host = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(LAG + 1)]
        loss_ev = [torch.cuda.Event() for _ in range(LAG + 1)]
        seen = []

        def prefetch(i):
            with torch.cuda.stream(copy_st):
                copy_st.wait_event(ev_free[i & 1])
                bufs[i & 1].copy_(host["rgb"], non_blocking=True)         # H2D of step i's targets (this rank's shard)
                ev_copied[i & 1].record(copy_st)

        def e2e_step(i):
            st.wait_event(ev_copied[i & 1])
            loc.rgb = bufs[i & 1]
            if not os.environ.get("MARF_BENCH_NO_BUMP"):    # (diagnostic switch: treat the copied targets as unchanged data)
                g.engine.bump_data_version()                # genuinely new data every step: cached derived inputs are rebuilt
            prefetch(i + 1)
            loss = m.train_iteration(var, None)
            ev_free[i & 1].record(st)
            if opt.warp.fix_first and m.fused_tail is None:
                g.warp_param.weight.data[0] = 0
            k = i % (LAG + 1)
            loss_host[k].copy_(loss.all.detach().reshape(1), non_blocking=True)       # D2H of the step's result
            loss_ev[k].record(st)
            if i >= LAG:
                j = (i - LAG) % (LAG + 1)
                loss_ev[j].synchronize()
                seen.append(float(loss_host[j]))

        prefetch(0)
        for i in range(3):
            e2e_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            e2e_step(i + 3)
        e1.record(st)
        barrier()
        assert all(v == v for v in seen) or os.environ.get("MARF_CHAIN_DBG") or os.environ.get("MARF_BWD_DBG"), "non-finite loss in the e2e arm"
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t) / steps
        res["e2e"] = dict(value=n_px_total / (e2e_ms * 1e-3), unit="pixel-samples/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                          ms_per_step=e2e_ms,
                          what="Model.train_iteration with --fused_optimizer (fused step + device Adam + fix_first); every step's targets are "
                               "copied from pinned host memory (prefetched on a copy stream during the previous step, double-buffered, "
                               "data version bumped so cached derived inputs are rebuilt) and the loss is read back every step (consumed "
                               "two steps later)")
    if sampler is not None and rank == 0:
        res["clocks"] = sampler.stop()                       # sampled over both timed loops (resident + e2e)

    # ---------------- per-kernel pass (roofline): the same resident-input steps again with CUDA-event pairs recorded on
    # the launching stream around every launch of the tensor-core kernel classes (marf_profile).  Separate from
    # the headline loop because an event between two launches suspends programmatic dependent launch there.
    kern = {}
    if kernels and precision == "bf16":
        ksteps = max(1, min(steps, 20))
        g.engine.profile(True)
        g.forward(var, mode="train")
        torch.cuda.synchronize()
        g.engine.profile_read()
        for _ in range(ksteps):
            flush.zero_()
            g.forward(var, mode="train")
        torch.cuda.synchronize()
        kern = {k: dict(us_per_launch=1e3 * ms / max(n, 1), launches_per_step=n / ksteps) for k, (ms, n) in g.engine.profile_read().items() if n}
        g.engine.profile(False)
    res["_kern"], res["_wl"] = kern, wl
    g.engine.close()
    del m, g, var
    torch.cuda.empty_cache()
    return res


def run_marf(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    ctx = dict(world=world, rank=rank, device=device,
               flush=torch.empty(256 << 20, dtype=torch.uint8, device=device))       # > 126 MB L2
    sampler = ClockSampler(local)
    head = measure(args.workload, args.precision, args, ctx, args.steps, args.warmup, sampler=sampler)
    wl, kern = head.pop("_wl"), head.pop("_kern")
    pk = peaks()
    line = dict(metric="pixel-samples/sec (fwd+bwd+warp grad)", value=head["value"], unit="pixel-samples/s", n_gpus=args.gpus,
                steps=args.steps, warmup=max(3, args.warmup), ms_per_step=head["ms_per_step"], higher_is_better=True,
                scaling=wl["scaling"], vs_baseline=None, dtype="f32" if args.precision == "fp32" else "bf16 (fp32 accumulate)",
                data="synthetic",
                config=dict(workload=wl["name"], precision=args.precision, pixel_samples_per_step=head["pixel_samples_per_step"],
                            flop_per_pixel_sample=wl["flop"], l2="flushed between timed steps (256 MiB memset)",
                            timing="sum of per-step CUDA-event intervals on the launch stream, max over ranks",
                            wall_s_timed_loop=head["wall_s_timed_loop"]),
                clocks=head.get("clocks"), e2e=head["e2e"], gpu_launches=head["gpu_launches"])
    line.update(roofline_lines(wl, args.workload, args.precision, kern, pk, head["pixel_samples_per_step"] // world, head["tflops"], args.gpus))
    # ---------------- the other named configurations and the fp32 parity mode, same run (N = 1 only: they are single-GPU lines)
    if args.gpus == 1 and not args.headline_only:
        short = max(3, min(args.steps, 10))
        for extra in ("config2", "config4", "config5"):
            if extra == args.workload:
                continue
            try:
                r = measure(extra, args.precision, args, ctx, short if extra != "config2" else max(100, args.steps), 3, e2e=True, kernels=True)   # (config 2: sub-ms steps)
                ewl, ekern = r.pop("_wl"), r.pop("_kern")
                r.update(roofline_lines(ewl, extra, args.precision, ekern, pk, r["pixel_samples_per_step"], r["tflops"], 1))
                r.pop("kernels", None)
                line[extra] = r
            except Exception as ex:
                line[extra] = dict(unavailable=repr(ex)[:200])
        # the fp32 parity mode last (its tensor-core GEMMs run the GPU into the power cap: the sub-millisecond bf16 lines above are
        # measured before it)
        if args.precision == "bf16":
            try:
                f = measure(args.workload, "fp32", args, ctx, max(2, min(args.steps, 3)), 1, e2e=False, kernels=False)
                line["value_fp32"] = f["value"]
                line["fp32"] = dict(value=f["value"], unit="pixel-samples/s", ms_per_step=f["ms_per_step"], steps=f["steps"],
                                    what="the same workload with precision=fp32 (the <=1e-3 parity mode: 3xTF32 tensor-core GEMMs for the wide layers, fp32 activations), resident inputs")
            except Exception as ex:                      # never lose the headline line to an extra
                line["fp32"] = dict(unavailable=repr(ex)[:200])
        if args.precision == "bf16" and args.workload != "config2" and isinstance(line.get("config2"), dict) and "value" in line["config2"]:
            try:                                         # ... and on the reference's own default configuration
                f2 = measure("config2", "fp32", args, ctx, 20, 3, e2e=True, kernels=False)
                line["config2"].update(value_fp32=f2["value"], ms_per_step_fp32=f2["ms_per_step"], e2e_fp32=f2["e2e"])
            except Exception as ex:
                line["config2"]["fp32_unavailable"] = repr(ex)[:200]
    if rank == 0:
        if args.gpus == 1 and not args.no_cpu:
            base, _ = cpu_reference(wl, steps=3, warmup=1, patches=cpu_sample_patches(wl))
            line["cpu_baseline"] = dict(value=base["value"], unit="pixel-samples/s", cores=base["cores"], kind="port",
                                        sample=base["sample"])
            try:                                                         # informational leg: never fail the bench line
                line["eager_gpu_baseline"] = eager_gpu_reference(workload("config2", 1), device)
            except Exception as ex:
                line["eager_gpu_baseline"] = dict(unavailable=repr(ex)[:200])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="marf", choices=["marf", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MARF_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--workload", default=os.environ.get("MARF_BENCH_WORKLOAD", "config4"))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--headline-only", action="store_true", help="skip the extra configurations / fp32 leg")
    args = ap.parse_args()
    os.chdir(ROOT)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_marf(args)


if __name__ == "__main__":
    main()
