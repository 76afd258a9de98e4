# ncu --set full of the K2 kernel at C2 for the current and the previous library; summaries into gpurun_out/r02e
OUT=gpurun_out/r02e; mkdir -p $OUT
PREV=$PWD/rag-faiss-embedding_b200/lib/libb200flat_prev.so
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-series --no-c4 --no-parity"
ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $OUT/prof_k2_c2_new -f $CMD > $OUT/ncu_new.log 2>&1
B200FLAT_LIB=$PREV ncu --set full --import-source on --clock-control none -k regex:tensor_scan --launch-skip 4 -c 1 -o $OUT/prof_k2_c2_prev -f $CMD > $OUT/ncu_prev.log 2>&1
python tools/ncu_summary.py $OUT/prof_k2_c2_new.ncu-rep "new" > $OUT/k2_c2_new.txt
python tools/ncu_summary.py $OUT/prof_k2_c2_prev.ncu-rep "prev" > $OUT/k2_c2_prev.txt
paste -d'|' $OUT/k2_c2_new.txt $OUT/k2_c2_prev.txt | head -80
ls -la $OUT
