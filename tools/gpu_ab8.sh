# same-box A/B of two builds of the library: lib/libb200flat_ab.so (before) against lib/libb200flat.so (after)
OUT=gpurun_out/r02t; mkdir -p $OUT
run() { name=$1; wl=$2; extra=$3; shift 3
  env "$@" timeout 400 python bench.py --workload $wl $extra --no-series --no-c4 --no-cpu-baseline --no-parity --steps 20 --warmup 5 > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1]); e=j["engine"]; r=j["roofline"]
    print("$name", "step", j["ms_per_step"], "kernel", r["kernel_ms"], "x", r["launches_per_search"], "frac", r["frac"], "pipe", r["pipeline_ms"], "surv", e.get("filter_survivors_per_query"), "fb", e.get("fallback_queries"))
except Exception as ex:
    print("$name FAILED", ex)
PY
}
AB=$PWD/rag-faiss-embedding_b200/lib/libb200flat_ab.so
for rep in 1 2; do
for wl in $WLS; do
  set -- $(echo $wl | tr ':' ' '); w=$1; nq=$2
  ex=""; [ -n "$nq" ] && ex="--nq $nq"
  run ${w}${nq}_before_$rep $w "$ex" B200FLAT_LIB=$AB
  run ${w}${nq}_after_$rep $w "$ex" X=1
done
done
