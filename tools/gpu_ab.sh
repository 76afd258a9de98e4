#!/bin/bash
# Same-box A/B of the previous (lib/libb200flat_prev.so) and the current kernel library on a list of bench workloads.
#   bash tools/gpu_ab.sh <outdir> wl1 wl2 ...
OUT=$1; shift
mkdir -p $OUT
PREV=$PWD/rag-faiss-embedding_b200/lib/libb200flat_prev.so
for wl in "$@"; do
  for lib in new prev; do
    if [ $lib = prev ]; then L="B200FLAT_LIB=$PREV"; else L="X=1"; fi
    env $L timeout 400 python bench.py --workload $wl --no-series --no-c4 --no-cpu-baseline --steps 20 --warmup 5 > $OUT/ab_${wl}_${lib}.json 2> $OUT/ab_${wl}_${lib}.err
    python - <<PY
import json
try:
    j=json.loads(open("$OUT/ab_${wl}_${lib}.json").read().strip().splitlines()[-1])
    e=j.get("engine") or j["config"]
    print("$wl $lib step", j["ms_per_step"], "kernel", j["roofline"]["kernel_ms"], "x", j["roofline"]["launches_per_search"], "frac", j["roofline"]["frac"], "pipe", j["roofline"]["pipeline_ms"], "surv", e.get("filter_survivors_per_query"), "fb", e.get("fallback_queries"), "par", (j.get("parity_check") or {}).get("recall"), "clk", (j.get("clocks") or {}).get("sm_mhz"), (j.get("clocks") or {}).get("reasons"))
except Exception as ex:
    print("$wl $lib FAILED", ex)
PY
  done
done
