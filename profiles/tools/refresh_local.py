"""Turn the raw captures of profiles/tools/refresh_gpu.sh (gpurun_out/) into the tracked artifacts under profiles/:
r01_launches_bf16_config2.csv (raw launch list), r01_step_kernels_bf16_config2.csv (per-kernel summary of one step),
r01_top_kernels_full.txt (selected metrics of the --set full capture), r01_kernel_traffic.json (DRAM bytes per launch, read by
bench.py for roofline.traffic), r01_bench_config2.json (the bench line)."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "launches_final.csv"), os.path.join(P, "r01_launches_bf16_config2.csv"))
shutil.copy(os.path.join(G, "bench_final.json"), os.path.join(P, "r01_bench_config2.json"))
summ = subprocess.run([sys.executable, os.path.join(P, "tools", "summarize_launches.py"), os.path.join(G, "launches_final.csv")],
                      capture_output=True, text=True, check=True).stdout
open(os.path.join(P, "r01_step_kernels_bf16_config2.csv"), "w").write(summ)
raw = subprocess.run(["ncu", "-i", os.path.join(G, "top_kernels_final.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_subpipe_hmma.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__cluster_dim_x', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']
mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
out = ["ncu --set full --import-source on --clock-control none, the three dominant kernels of ONE training step (config2, bf16), B200,",
       "round-1 final code; report: gpurun_out/top_kernels_final.ncu-rep (not tracked).  Times are cold-cache and serialised",
       "(bench.py kernels[] has the in-step CUDA-event times)."]
traffic = {}
names = {"k_tc_chain<0": "k_tc_chain<fwd>", "k_tc_chain<1": "k_tc_chain<dx>", "k_tc_dw": "k_tc_dw"}
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    out += ["", "==  " + name]
    d = {}
    for i, h in enumerate(hdr):
        if h in keep:
            out.append('   %-80s %s %s' % (h, r[i], units[i]))
            d[h] = (float(r[i].replace(",", "")), units[i])
    b = lambda k: d[k][0] * mult[d[k][1]]
    for key, nice in names.items():
        if key in name:
            traffic[nice] = int(b('dram__bytes_read.sum') + b('dram__bytes_write.sum'))
for line in summ.splitlines():
    if "k_tc_gemm<64, 3" in line:
        f = line.split(",")
        traffic["k_tc_gemm<64,warp_grad>"] = int((float(f[-5]) + float(f[-4])) * 1e6)
open(os.path.join(P, "r01_top_kernels_full.txt"), "w").write("\n".join(out) + "\n")
json.dump({"workload": "config2", "precision": "bf16",
           "source": "ncu --set full --clock-control none, one capture of one training step (profiles/r01_top_kernels_full.txt; the "
                     "warp-grad GEMM from the --metrics launch list); dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "dram_bytes_per_launch": traffic}, open(os.path.join(P, "r01_kernel_traffic.json"), "w"), indent=1)
print(summ)
print(json.dumps(traffic))
