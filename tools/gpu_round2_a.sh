#!/bin/bash
# One gpurun call: smoke the new K2, the GPU test tier, the default bench line, same-box A/B of the previous and the
# current kernel library on the shapes the work split matters for, and the C1 CLI acceptance.
OUT=gpurun_out/r02a
mkdir -p $OUT
PREV=$PWD/rag-faiss-embedding_b200/lib/libb200flat_prev.so
if timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; then echo "smoke ok"; else echo "SMOKE FAILED (new library)"; tail -5 $OUT/smoke.log; export B200FLAT_LIB=$PREV; fi
(timeout 1300 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log)
tail -6 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?"; tail -c 400 $OUT/bench_default.err
for wl in c2 c2_nq4096 c2_shard8 c2_shard4 c2_nq32 c2_nq128; do
  for lib in new prev; do
    if [ $lib = prev ]; then L="B200FLAT_LIB=$PREV"; else L="X=1"; fi
    env $L timeout 300 python bench.py --workload $wl --no-series --no-c4 --no-cpu-baseline --steps 20 --warmup 5 > $OUT/ab_${wl}_${lib}.json 2> $OUT/ab_${wl}_${lib}.err
    python - <<PY
import json
try:
    j=json.loads(open("$OUT/ab_${wl}_${lib}.json").read().strip().splitlines()[-1])
    print("$wl $lib", j["ms_per_step"], "kernel", j["roofline"]["kernel_ms"], "frac", j["roofline"]["frac"], "pipe", j["roofline"]["pipeline_ms"], "surv", j["engine"]["filter_survivors_per_query"] if "engine" in j else j["config"].get("filter_survivors_per_query"), "fb", (j.get("engine") or j["config"]).get("fallback_queries"), "par", (j.get("parity_check") or {}).get("recall"))
except Exception as e:
    print("$wl $lib FAILED", e)
PY
  done
done
timeout 300 python tools/c1_reference_cli.py > $OUT/c1_cli.json 2> $OUT/c1_cli.err; echo "c1 exit $?"; head -c 900 $OUT/c1_cli.json
