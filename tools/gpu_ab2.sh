OUT=gpurun_out/r02c; mkdir -p $OUT
run() { # name, env..., workload
  name=$1; wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --no-series --no-c4 --no-cpu-baseline --no-parity --steps 20 --warmup 5 > $OUT/$name.json 2> $OUT/$name.err
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
run c2_new c2 X=1
run c2_prev c2 B200FLAT_LIB=$PREV
run c2_new_die0 c2 B200FLAT_DIE_MODE=0
run c2_prev_die0 c2 B200FLAT_LIB=$PREV B200FLAT_DIE_MODE=0
run c2_new_r19 c2 B200FLAT_ROUND_TILES=19
run c2_new_r74 c2 B200FLAT_ROUND_TILES=74
run c2_new_r148 c2 B200FLAT_ROUND_TILES=148
run c2_new_r19_die0 c2 B200FLAT_ROUND_TILES=19 B200FLAT_DIE_MODE=0
run c2_new_nopair c2 B200FLAT_NO_PAIR=1
run c2_prev_nopair c2 B200FLAT_LIB=$PREV B200FLAT_NO_PAIR=1
run c2_new_stage c2 B200FLAT_STAGE1=-1
run s8_new c2_shard8 X=1
run s8_prev c2_shard8 B200FLAT_LIB=$PREV
run s8_new_die0 c2_shard8 B200FLAT_DIE_MODE=0
run s8_new_r3 c2_shard8 B200FLAT_ROUND_TILES=3
run s8_new_r12 c2_shard8 B200FLAT_ROUND_TILES=12
run nq128_new c2_nq128 X=1
run nq128_prev c2_nq128 B200FLAT_LIB=$PREV
run nq256_new c2_nq128 X=1
