export OUT=gpurun_out/r02i; mkdir -p $OUT
bash tools/gpu_ab4.sh 2>&1 | grep -E "^c2_new|^c2_prev|shard8|nq4096|nq512"
EXTRA=1
PREV=$PWD/rag-faiss-embedding_b200/lib/libb200flat_prev.so
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-series --no-c4 --no-parity"
ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $OUT/prof_k2_c2_new -f $CMD > $OUT/ncu_new.log 2>&1
B200FLAT_LIB=$PREV ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $OUT/prof_k2_c2_prev -f $CMD > $OUT/ncu_prev.log 2>&1
python tools/ncu_summary.py $OUT/prof_k2_c2_new.ncu-rep "new" > $OUT/k2_c2_new.txt
python tools/ncu_summary.py $OUT/prof_k2_c2_prev.ncu-rep "prev" > $OUT/k2_c2_prev.txt
grep -E "inst_executed|cycles_elapsed.avg |time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|stalled_wait|stalled_long|no_instruction|dram__bytes_read.sum " $OUT/k2_c2_new.txt $OUT/k2_c2_prev.txt
