OUT=gpurun_out/r02d; mkdir -p $OUT
run() { # name, workload, extra bench args (quoted), env...
  name=$1; wl=$2; extra=$3; shift 3
  env "$@" timeout 300 python bench.py --workload $wl $extra --no-series --no-c4 --no-cpu-baseline --no-parity --steps 20 --warmup 5 > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1]); e=j["engine"]
    print("$name", "step", j["ms_per_step"], "kernel", j["roofline"]["kernel_ms"], "x", j["roofline"]["launches_per_search"], "pipe", j["roofline"]["pipeline_ms"], "surv", e.get("filter_survivors_per_query"), "fb", e.get("fallback_queries"))
except Exception as ex:
    print("$name FAILED", ex)
PY
}
PREV=$PWD/rag-faiss-embedding_b200/lib/libb200flat_prev.so
for nq in 256 512 1024 1280 2048; do
  run nq${nq}_new c2 "--nq $nq" X=1
  run nq${nq}_prev c2 "--nq $nq" B200FLAT_LIB=$PREV
done
run nq256_new_nopair c2 "--nq 256" B200FLAT_NO_PAIR=1
run nq256_prev_nopair c2 "--nq 256" B200FLAT_LIB=$PREV B200FLAT_NO_PAIR=1
run nq512_new_nopair c2 "--nq 512" B200FLAT_NO_PAIR=1
run nq512_prev_nopair c2 "--nq 512" B200FLAT_LIB=$PREV B200FLAT_NO_PAIR=1
