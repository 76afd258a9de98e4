#!/bin/bash
# Runs on an N-GPU box (gpurun --gpus N): the multi-GPU test tier, the bench at N (and its reference arm), optionally C5.
#   bash tools/gpu_multi.sh <N> [c5]
N=$1; OUT=gpurun_out/r02s_n$N; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
(timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $OUT/pytest_sharded.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_sharded.log); tail -4 $OUT/pytest_sharded.log
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N exit $?"; tail -c 300 $OUT/bench_n$N.err
python - <<PY
import json
try:
    j=json.loads([l for l in open("$OUT/bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N value", j["value"], "ms", j["ms_per_step"], "kernel", j["roofline"]["kernel_ms"], "pipe", j["roofline"]["pipeline_ms"], "e2e", j["e2e"]["value"], "parity", j.get("parity_check"), "launches", j["gpu_launches"])
    c4=j.get("c4") or {}
    print("  c4", c4.get("value"), c4.get("ms_per_step"), (c4.get("roofline") or {}).get("frac"), c4.get("parity_check"), [ (s["nq"], s["value"], s["roofline"]["frac"]) for s in c4.get("small_batches",[])])
    print("  cpu", j.get("cpu_baseline"))
except Exception as e:
    print("parse failed", e)
PY
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err; echo "ref exit $?"; head -c 400 $OUT/bench_ref_n$N.json; echo
if [ "$2" = "c5" ]; then
  timeout 1200 $TR tools/c5_pipeline.py --chunks 5000000 --queries 10000 > $OUT/c5_n$N.json 2> $OUT/c5_n$N.err; echo "c5 exit $?"; tail -1 $OUT/c5_n$N.json | head -c 1200; echo; tail -c 300 $OUT/c5_n$N.err
fi
