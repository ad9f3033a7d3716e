# usage: bash profiles/tools/bwd_split_probe.sh "ENV1=a ENV2=b" "ENV1=c" ...   (one bench run per argument, with a k_tc_bwd trace summary)
for e in "$@"; do
  env $e MARF_BWD_TRACE=12 python bench.py --workload ${WL:-config2} --steps 20 --warmup 3 --no-cpu --headline-only 2> gpurun_out/probe.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    ks={k['kernel']:round(k['us_per_launch'],1) for k in d.get('kernels',[])}
    print('$e | ms/step %.3f  value %.1f M |' % (d['ms_per_step'], d['value']/1e6), ks)
except Exception as ex:
    print('$e FAILED', ex)"
  grep "k_tc_bwd trace" gpurun_out/probe.err
  grep -i "error\|trap\|illegal" gpurun_out/probe.err | head -3
done
