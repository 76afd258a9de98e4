#!/bin/bash
# ncu --set full of the small-batch kernels on one shard of BASELINE configs[3] (12.5M x 384 bf16): the north star's HBM figure.
O=gpurun_out/r02y; mkdir -p $O
for nq in 1 32; do
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-series --no-c4 --no-parity --workload c4shard_nq$nq"
  timeout 400 $CMD > $O/bench_c4shard_nq$nq.json 2> $O/bench_c4shard_nq$nq.err || { echo "plain run failed"; tail -3 $O/bench_c4shard_nq$nq.err; exit 1; }
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:"tensor_scan|scan_kernel" --launch-skip 4 -c 1 -o $O/prof_c4shard_nq$nq -f $CMD > $O/ncu_c4shard_nq$nq.log 2>&1
  python tools/ncu_summary.py $O/prof_c4shard_nq$nq.ncu-rep "ncu --set full --clock-control none, c4shard_nq$nq; 1 launch of: $CMD" > $O/r02_ncu_full_c4shard_nq$nq.txt 2>/dev/null
  grep "^kernel\|time_duration\|dram__bytes_read" $O/r02_ncu_full_c4shard_nq$nq.txt | cut -c1-150
  python -c "
import json; j=json.loads(open('$O/bench_c4shard_nq$nq.json').read().strip().splitlines()[-1]); print('live', j['roofline'])"
done
