#!/bin/bash
# ncu --set full of K2 on BASELINE configs[2] (10M x 768 IP, k = 100, batch 4096).
O=gpurun_out/r02z; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-series --no-c4 --no-parity --workload c3_nq4096"
timeout 500 $CMD > $O/bench_c3.json 2> $O/bench_c3.err || { echo "plain run failed"; tail -3 $O/bench_c3.err; exit 1; }
timeout 700 ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 8 -c 1 -o $O/prof_k2_c3 -f $CMD > $O/ncu_k2_c3.log 2>&1
python tools/ncu_summary.py $O/prof_k2_c3.ncu-rep "ncu --set full --clock-control none, k2_c3_nq4096; 1 launch of: $CMD" > $O/r02_ncu_full_k2_c3_nq4096.txt 2>/dev/null
grep "^kernel\|time_duration\|dram__bytes_read.sum \[\|pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" $O/r02_ncu_full_k2_c3_nq4096.txt | cut -c1-160
python -c "
import json; j=json.loads(open('$O/bench_c3.json').read().strip().splitlines()[-1]); print('live', j['roofline'], j['ms_per_step'])"
