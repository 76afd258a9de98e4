# same-box A/B of several builds of the library: VARIANTS="ab expa expb" name lib/libb200flat_<v>.so; "product" is lib/libb200flat.so
OUT=gpurun_out/r02u; mkdir -p $OUT
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
for rep in 1 2; do
for wl in $WLS; do
  set -- $(echo $wl | tr ':' ' '); w=$1; nq=$2
  ex=""; [ -n "$nq" ] && ex="--nq $nq"
  for v in $VARIANTS; do
    if [ $v = product ]; then run ${w}${nq}_${v}_$rep $w "$ex" X=1; else run ${w}${nq}_${v}_$rep $w "$ex" B200FLAT_LIB=$PWD/rag-faiss-embedding_b200/lib/libb200flat_$v.so; fi
  done
done
done
