#!/bin/bash
# A/B of library builds on the bench workload: tools/ab_bench.sh base v1 v2 ...  (lib/libsprsolve_b200_<name>.so; "new" = the default lib)
for L in "$@"; do
  if [ "$L" = new ]; then unset SPB_LIB; else export SPB_LIB=$PWD/sprsolve_b200/lib/libsprsolve_b200_$L.so; fi
  python bench.py --no-cpu --no-configs --no-full-solve --steps 3 --warmup 3 > gpurun_out/ab_$L.json 2> gpurun_out/ab_$L.err
  python - "$L" <<'P'
import json, sys
L = sys.argv[1]
d = json.loads(open(f"gpurun_out/ab_{L}.json").read().strip().splitlines()[-1])
r = d["roofline"]
print(L, "%.2f it/s" % d["value"], "spmv %.3f ms" % r["avg_launch_ms"], "frac %.3f" % r["frac"], "c2 %.4f ms" % d["spmv_c2"]["ms"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], flush=True)
P
done
