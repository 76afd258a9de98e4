OUT=${OUT:-gpurun_out/r02f}; mkdir -p $OUT
run() { # name, workload, extra bench args (quoted), env...
  name=$1; wl=$2; extra=$3; shift 3
  env "$@" timeout 400 python bench.py --workload $wl $extra --no-series --no-c4 --no-cpu-baseline --no-parity --steps 20 --warmup 5 > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1]); e=j["engine"]
    print("$name", "step", j["ms_per_step"], "kernel", j["roofline"]["kernel_ms"], "x", j["roofline"]["launches_per_search"], "pipe", j["roofline"]["pipeline_ms"], "surv", e.get("filter_survivors_per_query"), "fb", e.get("fallback_queries"), "resc", e.get("rescued_queries"))
except Exception as ex:
    print("$name FAILED", ex)
PY
}
PREV=$PWD/rag-faiss-embedding_b200/lib/libb200flat_prev.so
for spec in "c2:" "c2:--nq 256" "c2:--nq 512" "c2_nq4096:" "c2_shard8:" "c2_shard4:" "c2_nq32:" "c2_nq128:" ${EXTRA_SPECS}; do
  wl=${spec%%:*}; extra=${spec#*:}; tag=$(echo "$wl$extra" | tr -d ' -')
  run ${tag}_new $wl "$extra" X=1
  run ${tag}_prev $wl "$extra" B200FLAT_LIB=$PREV
done
