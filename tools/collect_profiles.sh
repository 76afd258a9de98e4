#!/bin/bash
# Runs on the GPU box (under gpurun): the launch list of the default bench command and the ncu --set full captures that
# profiles/ summarises (each only after the same command exited 0 without ncu).  Everything lands in gpurun_out/r02p.
O=gpurun_out/r02p; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-series --no-c4"
timeout 300 $CMD > $O/bench_short.json 2> $O/bench_short.err || { echo "plain run failed"; tail -3 $O/bench_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2_default.csv $CMD > $O/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $O/prof_k2_c2 -f $CMD > $O/ncu_k2_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $O/prof_k2_nq32 -f $CMD --workload c2_nq32 > $O/ncu_k2_nq32.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $O/prof_k2_nq4096 -f $CMD --workload c2_nq4096 > $O/ncu_k2_nq4096.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:scan_kernel --launch-skip 4 -c 1 -o $O/prof_k1_nq1 -f $CMD --workload c2_nq1 > $O/ncu_k1_nq1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:merge_lists --launch-skip 4 -c 1 -o $O/prof_merge_c2 -f $CMD > $O/ncu_merge_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:merge_lists --launch-skip 4 -c 1 -o $O/prof_merge_shard8 -f $CMD --workload c2_shard8 > $O/ncu_merge_shard8.log 2>&1
for r in k2_c2 k2_nq32 k2_nq4096 k1_nq1 merge_c2 merge_shard8; do
  python tools/ncu_summary.py $O/prof_$r.ncu-rep "ncu --set full --clock-control none, $r; 1 launch of: $CMD" > $O/r02_ncu_full_$r.txt 2>/dev/null
done
ls -la $O | head -30
