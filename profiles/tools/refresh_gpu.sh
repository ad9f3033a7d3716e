#!/bin/bash
# Runs ON the GPU box (gpurun -- 'bash profiles/tools/refresh_gpu.sh'): regenerates the raw inputs of the profile artifacts
# under gpurun_out/ — a bench line, the ncu launch list of one step and a --set full capture of the three dominant kernels.
# profiles/tools/refresh_local.py then turns them into the tracked files under profiles/.
set -u
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_final.json
ncu --kernel-name regex:k_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 120 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:"k_tc_dw|k_tc_chain" --launch-skip 9 --launch-count 3 \
    -o gpurun_out/top_kernels_final python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -c 300 gpurun_out/bench_final.json
