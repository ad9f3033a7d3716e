#!/bin/bash
# Runs ON the GPU box (gpurun -- 'bash profiles/tools/refresh_gpu.sh'): regenerates the raw inputs of the round-2 profile
# artifacts under gpurun_out/ — the default bench line, the ncu launch lists of one config-2 step and of config-4 chunks, and a
# --set full capture of the dominant kernels.  profiles/tools/refresh_local.py then turns them into the tracked files under profiles/.
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
C2="python bench.py --workload config2 --steps 2 --warmup 3 --no-cpu --headline-only"
$C2 > gpurun_out/plain_c2.log 2>&1 && ncu --kernel-name regex:k_ --metrics $M --clock-control none -c 150 --csv --log-file gpurun_out/launches_r02_config2.csv $C2 > gpurun_out/ncu_c2.log 2>&1
C4="python bench.py --workload config4 --steps 1 --warmup 3 --no-cpu --headline-only"
$C4 > gpurun_out/plain_c4.log 2>&1 && ncu --kernel-name regex:k_ --metrics $M --clock-control none -s 1400 -c 120 --csv --log-file gpurun_out/launches_r02_config4.csv $C4 > gpurun_out/ncu_c4.log 2>&1
$C2 > gpurun_out/plain_c2b.log 2>&1 && ncu --set full --import-source on --clock-control none --kernel-name regex:"k_tc_bwd|k_tc_chain" --launch-skip 8 --launch-count 2 -o gpurun_out/r02_top_kernels $C2 > gpurun_out/ncu_full.log 2>&1
F2="python bench.py --precision fp32 --workload config2 --steps 2 --warmup 3 --no-cpu --headline-only"
$F2 > gpurun_out/plain_f2.log 2>&1 && ncu --kernel-name regex:k_ --metrics $M --clock-control none --launch-skip 700 -c 340 --csv --log-file gpurun_out/launches_r02_fp32_config2.csv $F2 > gpurun_out/ncu_f2.log 2>&1
tail -c 400 gpurun_out/bench_r02_final.json
