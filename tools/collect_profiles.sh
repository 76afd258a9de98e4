#!/bin/bash
# Runs on the GPU box (under gpurun): the bench series, the launch list of the default bench command and the
# ncu --set full captures that profiles/ summarises.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
timeout 600 python bench.py > $O/bench_default.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.log 2>&1
for W in c2_nq1 c2_nq32 c2_nq128 c2_nq4096 c4shard_nq1 c4shard_nq32; do timeout 300 $B --workload $W > $O/bench_$W.log 2>&1; done
timeout 300 $B --workload c2_nq1 --algo tensor > $O/bench_c2_nq1_tensor.log 2>&1
S="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for W in c3_nq1 c3_nq32 c3_nq4096 c4shard; do timeout 400 $S --workload $W > $O/bench_$W.log 2>&1; done
timeout 300 $S --workload c3_nq1 --algo tensor > $O/bench_c3_nq1_tensor.log 2>&1
python tools/summarize_bench.py $O/bench_default.log $O/bench_c2_*.log $O/bench_c3_*.log $O/bench_c4shard*.log
tail -1 $O/bench_reference.log | cut -c1-300
# launch list of the default bench command (after it exited 0 without ncu above)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv $CMD > $O/ncu_launches.log 2>&1
# one full capture per dominant kernel
ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $O/prof_k2_c2 -f $CMD > $O/ncu_k2_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $O/prof_k2_nq32 -f $CMD --workload c2_nq32 > $O/ncu_k2_nq32.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:scan_kernel --launch-skip 4 -c 1 -o $O/prof_k1_nq1 -f $CMD --workload c2_nq1 > $O/ncu_k1_nq1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:merge_lists --launch-skip 4 -c 1 -o $O/prof_merge_c2 -f $CMD > $O/ncu_merge_c2.log 2>&1
ncu --set full --clock-control none -k regex:ingest_kernel -c 1 -o $O/prof_k5_ingest -f python tools/ingest_bench.py > $O/ncu_k5.log 2>&1
ls -la $O/*.ncu-rep
