#!/bin/bash
# ncu --set full of K2 on one shard of BASELINE configs[3] (12.5M x 384 bf16, batch 4096): the north star's tensor-pipe figure.
O=gpurun_out/r02x; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-series --no-c4 --no-parity --workload c4shard"
timeout 400 $CMD > $O/bench_c4shard.json 2> $O/bench_c4shard.err || { echo "plain run failed"; tail -3 $O/bench_c4shard.err; exit 1; }
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $O/prof_k2_c4shard -f $CMD > $O/ncu_k2_c4shard.log 2>&1
python tools/ncu_summary.py $O/prof_k2_c4shard.ncu-rep "ncu --set full --clock-control none, k2_c4shard; 1 launch of: $CMD" > $O/r02_ncu_full_k2_c4shard.txt 2>/dev/null
grep "time_duration\|tensor\|dram__bytes_read.sum \[" $O/r02_ncu_full_k2_c4shard.txt | head
