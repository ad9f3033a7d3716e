"""Summarise an `ncu --csv --metrics ...` launch list: per kernel launches / time / DRAM bytes for the LAST full step."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; rows = rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows:
    key = (r[ix["ID"]], r[ix["Kernel Name"]])
    per.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
launches = [(k[1], v) for k, v in per.items()]
# last step = from the last k_sl3_to_SL3 launch onward... take the last occurrence block
starts = [i for i, (n, _) in enumerate(launches) if "k_pack_table" in n]
# a step = [k_sl3_to_SL3 just before k_pack_table ... the launch before the next step's k_sl3_to_SL3]
if len(starts) >= 2: seg = launches[starts[-2] - 1:starts[-1] - 1]
else: seg = launches
agg = collections.OrderedDict()
for n, v in seg:
    n = n.split("(")[0]
    a = agg.setdefault(n, dict(n=0, us=0.0, rd=0.0, wr=0.0, tp=0.0, l2=0.0))
    a["n"] += 1; a["us"] += v.get("gpu__time_duration.sum", 0) / 1e3
    a["rd"] += v.get("dram__bytes_read.sum", 0); a["wr"] += v.get("dram__bytes_write.sum", 0)
    a["tp"] += v.get("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", 0)
    a["l2"] += v.get("lts__t_bytes.sum", 0)
tot = sum(a["us"] for a in agg.values())
print("kernel,launches,us,share_pct,dram_read_MB,dram_write_MB,dram_GB/s,l2_MB,tensor_pipe_pct")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f'{n},{a["n"]},{a["us"]:.1f},{100*a["us"]/tot:.1f},{a["rd"]/1e6:.1f},{a["wr"]/1e6:.1f},{(a["rd"]+a["wr"])/a["us"]/1e3:.0f},{a["l2"]/1e6:.0f},{a["tp"]/a["n"]:.1f}')
print(f'TOTAL,{sum(a["n"] for a in agg.values())},{tot:.1f},100,{sum(a["rd"] for a in agg.values())/1e6:.1f},{sum(a["wr"] for a in agg.values())/1e6:.1f},,,')
