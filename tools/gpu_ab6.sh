OUT=gpurun_out/r02l; mkdir -p $OUT
run() { name=$1; wl=$2; extra=$3; shift 3
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
run c3_new c3_nq4096 "" X=1
run c3_new_u144 c3_nq4096 "" B200FLAT_MAX_UNITS=144
run c3_prev c3_nq4096 "" B200FLAT_LIB=$PREV
run c2_new c2 "" X=1
run c2_new_u72 c2 "" B200FLAT_MAX_UNITS=72
run c2_prev c2 "" B200FLAT_LIB=$PREV
run c3_nq32_new c3_nq32 "" X=1
run c3_nq32_prev c3_nq32 "" B200FLAT_LIB=$PREV
run c3_nq1t_new c3_nq1 "--algo tensor" X=1
run c3_nq1t_prev c3_nq1 "--algo tensor" B200FLAT_LIB=$PREV
run c4_nq32_new c4shard_nq32 "" X=1
run c4_nq32_prev c4shard_nq32 "" B200FLAT_LIB=$PREV
run c4_nq1_new c4shard_nq1 "" X=1
run c4_nq1_prev c4shard_nq1 "" B200FLAT_LIB=$PREV
