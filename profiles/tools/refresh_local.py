"""Turn the raw captures of profiles/tools/refresh_gpu.sh (gpurun_out/) into the tracked round-2 artifacts under profiles/:
r02_launches_config{2,4}.csv (raw ncu launch lists), r02_step_kernels_config{2,4}.csv (per-kernel summary of one config-2 step /
one 2^20-pixel chunk of config 4), r02_top_kernels_full.txt (selected metrics of the --set full capture), r02_kernel_traffic.json
(DRAM bytes per launch, read by bench.py for roofline.traffic), r02_bench.json (the default bench line)."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr, rows = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault((r[ix["ID"]], r[ix["Kernel Name"]]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    return [(k[1].split("(")[0].replace("void ", "").replace("marf::", "").replace("tc::", ""), v) for k, v in per.items()]


def summary(seg, title):
    out = [title, "kernel,us,share_pct,dram_read_MB,dram_write_MB,dram_GB/s,l2_MB,l2_TB/s,l2_hit_pct,tensor_pipe_pct"]
    tot = sum(v.get("gpu__time_duration.sum", 0) for _, v in seg) / 1e3
    for n, v in seg:
        us = v.get("gpu__time_duration.sum", 0) / 1e3
        rd, wr, l2 = v.get("dram__bytes_read.sum", 0), v.get("dram__bytes_write.sum", 0), v.get("lts__t_bytes.sum", 0)
        out.append(f'{n},{us:.1f},{100 * us / tot:.1f},{rd / 1e6:.1f},{wr / 1e6:.1f},{(rd + wr) / us / 1e3:.0f},{l2 / 1e6:.0f},{l2 / us / 1e6:.1f},'
                   f'{v.get("lts__t_sector_hit_rate.pct", 0):.1f},{v.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 0):.1f}')
    rd = sum(v.get("dram__bytes_read.sum", 0) for _, v in seg)
    wr = sum(v.get("dram__bytes_write.sum", 0) for _, v in seg)
    out.append(f"TOTAL,{tot:.1f},100,{rd / 1e6:.1f},{wr / 1e6:.1f},{(rd + wr) / tot / 1e3:.0f},,,,")
    return "\n".join(out) + "\n", rd + wr


traffic = {}
for cfg in ("config2", "config4"):
    src = os.path.join(G, f"launches_r02_{cfg}.csv")
    shutil.copy(src, os.path.join(P, f"r02_launches_{cfg}.csv"))
    L = launches(src)
    if cfg == "config2":
        st = [i for i, (n, _) in enumerate(L) if "k_pack_table" in n]
        seg = L[st[-2]:st[-1]]
        title = "# one training step of config 2 (216,000 px-samples, image MLP + mask head + edge term), bf16; ncu --clock-control none (cold-cache, serialised: compare shares)"
    else:
        st = [i for i, (n, _) in enumerate(L) if "k_encode" in n]
        seg = L[st[2]:st[3]]
        title = "# one 2^20-pixel chunk of config 4 (64 chunks per step; image MLP, disk masks), bf16; ncu --clock-control none"
    txt, tot = summary(seg, title)
    open(os.path.join(P, f"r02_step_kernels_{cfg}.csv"), "w").write(txt)
    print(txt)
    names = {"k_tc_chain<0": "k_tc_chain<fwd>", "k_tc_chain<1": "k_tc_chain<dx>", "k_tc_bwd": "k_tc_bwd", "k_tc_dw": "k_tc_dw", "k_tc_gemm<64, 3": "k_tc_gemm<64,warp_grad>"}
    traffic[cfg] = {}
    for n, v in seg:
        for key, nice in names.items():
            if key in n:
                traffic[cfg][nice] = int(v.get("dram__bytes_read.sum", 0) + v.get("dram__bytes_write.sum", 0))
# precision=fp32, config 2: one step between two k_sl3_backward launches
src = os.path.join(G, "launches_r02_fp32_config2.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, "r02_launches_fp32_config2.csv"))
    L = launches(src)
    st = [i for i, (n, _) in enumerate(L) if "k_sl3_backward" in n]
    seg = L[st[-2] + 1:st[-1] + 1]
    agg = collections.OrderedDict()
    for n, v in seg:                      # (one line per kernel: ~100 launches per step)
        a = agg.setdefault(n, collections.Counter())
        a["launches"] += 1
        for k, x in v.items():
            a[k] += x
    rows_ = sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"])
    tot = sum(a["gpu__time_duration.sum"] for _, a in rows_) / 1e3
    out = ["# one training step of config 2 (216,000 px-samples, image MLP + mask head + edge term), precision=fp32 (3xTF32 tensor-core GEMMs), "
           "summed per kernel; ncu --clock-control none (cold-cache, serialised: compare shares)", "kernel,launches,us,share_pct,dram_read_MB,dram_write_MB,tensor_pipe_pct(avg)"]
    for n, a in rows_:
        us = a["gpu__time_duration.sum"] / 1e3
        out.append(f'{n},{a["launches"]},{us:.1f},{100 * us / tot:.1f},{a["dram__bytes_read.sum"] / 1e6:.1f},{a["dram__bytes_write.sum"] / 1e6:.1f},'
                   f'{a["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"] / a["launches"]:.1f}')
    out.append(f"TOTAL,{len(seg)},{tot:.1f},100,,,")
    open(os.path.join(P, "r02_step_kernels_fp32_config2.csv"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))

for cfg, t in traffic.items():
    json.dump({"workload": cfg, "precision": "bf16",
               "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one launch of each kernel inside a training step (profiles/r02_step_kernels_{cfg}.csv)",
               "dram_bytes_per_launch": t}, open(os.path.join(P, "r02_kernel_traffic.json" if cfg == "config4" else "r02_kernel_traffic_config2.json"), "w"), indent=1)
shutil.copy(os.path.join(G, "bench_r02_final.json"), os.path.join(P, "r02_bench.json"))

raw = subprocess.run(["ncu", "-i", os.path.join(G, "r02_top_kernels.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_hmma.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__cluster_dim_x', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']
out = ["ncu --set full --import-source on --clock-control none: the two dominant kernels of ONE training step (config 2, bf16), B200,",
       "round-2 final code; report: gpurun_out/r02_top_kernels.ncu-rep (not tracked).  Times are cold-cache and serialised",
       "(bench.py kernels[] has the in-step CUDA-event times)."]
for r in rows[2:]:
    out += ["", "==  " + r[hdr.index('Kernel Name')]]
    for i, h in enumerate(hdr):
        if h in keep:
            out.append('   %-80s %s %s' % (h, r[i], units[i]))
open(os.path.join(P, "r02_top_kernels_full.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-30:]))
