#!/bin/bash
# Round-end validation on one GPU: smoke, the GPU test tier, the default bench line and its reference arm, the ncu launch
# list of the bench command and one ncu --set full capture of the merge kernel.  Everything lands in gpurun_out/r02v.
O=gpurun_out/r02v; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
(timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log); tail -4 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-series --no-c4"
timeout 300 $CMD > $O/bench_short.json 2> $O/bench_short.err && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2_default.csv $CMD > $O/ncu_launches.log 2>&1
  ncu --set full --import-source on --clock-control none -k regex:merge_lists --launch-skip 4 -c 1 -o $O/prof_merge_c2 -f $CMD > $O/ncu_merge_c2.log 2>&1
  python tools/ncu_summary.py $O/prof_merge_c2.ncu-rep "ncu --set full --clock-control none, merge_c2; 1 launch of: $CMD" > $O/r02_ncu_full_merge_c2.txt 2>/dev/null
}
python - <<PY
import json
j=json.loads([l for l in open("$O/bench_default.json") if l.startswith("{")][-1]); r=j["roofline"]
print("C2 value", j["value"], "ms", j["ms_per_step"], "kernel", r["kernel_ms"], "frac", r["frac"], "e2e", j["e2e"]["value"], "launches", j["gpu_launches"], "parity", j.get("parity_check",{}).get("recall"))
for s in j.get("series",[]): print("  ", s.get("workload"), s.get("nq"), s.get("value"), s.get("roofline",{}).get("frac"), s.get("roofline",{}).get("bound"))
c4=j.get("c4") or {}; print("  c4", c4.get("value"), (c4.get("roofline") or {}).get("frac"), [(s["nq"], s["value"], s["roofline"]["frac"]) for s in c4.get("small_batches",[])])
print("  cpu", j.get("cpu_baseline"), "clocks", j.get("clocks"))
PY
