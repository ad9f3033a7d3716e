#!/bin/bash
# timing experiments on the fused backward launch (k_tc_bwd): role split, hand-over switches, per-pair trace
# usage (GPU box): bash profiles/tools/bwd_probe.sh [workload] [VAR=val ...]   (each VAR=val is one extra run)
WL=${1:-config2}
shift
run() {
  env "$@" python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu --headline-only 2> gpurun_out/probe.err | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
try:
    d=json.loads(txt[-1])
    ks={k['kernel']:round(k['us_per_launch'],1) for k in d.get('kernels',[])}
    print('$*', '| ms/step %.3f  value %.1f M  e2e %.1f M |' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6), ks)
except Exception as ex:
    print('$*', 'FAILED', ex)
"
  grep -A200 "k_tc_bwd trace" gpurun_out/probe.err | awk '{print}' | head -${TRACE_LINES:-90}
  grep -i "error\|trap\|illegal" gpurun_out/probe.err | head -3
}
run MARF_X=0
for v in "$@"; do run $v; done
