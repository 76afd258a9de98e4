OUT=gpurun_out/r02n; mkdir -p $OUT
run() { name=$1; wl=$2; extra=$3; shift 3
  env "$@" timeout 400 python bench.py --workload $wl $extra --no-series --no-c4 --no-cpu-baseline --no-parity --steps 20 --warmup 5 > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    j=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1]); e=j["engine"]; r=j["roofline"]
    print("$name", "step", j["ms_per_step"], "kernel", r["kernel_ms"], "x", r["launches_per_search"], "pipe", r["pipeline_ms"], "pipe-kernel", round(r["pipeline_ms"]-r["kernel_ms"]*r["launches_per_search"],4), "surv", e.get("filter_survivors_per_query"), "fb", e.get("fallback_queries"), "resc", e.get("rescued_queries"))
except Exception as ex:
    print("$name FAILED", ex)
PY
}
for wl in c2_shard8 c2_nq4096 c2_shard4 c4shard c3_nq4096; do
  run ${wl}_default $wl "" X=1
  run ${wl}_nowarp $wl "" B200FLAT_WARP_MERGE=0
  run ${wl}_nowarp_nostage $wl "" B200FLAT_WARP_MERGE=0 B200FLAT_STAGE1=-1
  run ${wl}_warp_nostage $wl "" B200FLAT_STAGE1=-1
done
run c2_warp c2 "" B200FLAT_WARP_MERGE=512
run c2_default c2 "" X=1
(timeout 1300 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log); tail -4 $OUT/pytest_gpu.log
