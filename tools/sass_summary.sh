#!/bin/bash
# SASS evidence that the hot kernels are tcgen05 / TMEM / TMA code: counts of the Blackwell mnemonics in the built
# library (cuobjdump -sass), per kernel.  Runs without a GPU.
#   bash tools/sass_summary.sh > profiles/r02_sass_summary.txt
LIB=${1:-rag-faiss-embedding_b200/lib/libb200flat.so}
echo "cuobjdump -sass $LIB ($(date -u +%Y-%m-%dT%H:%MZ)); nvcc $(nvcc --version | grep release | sed 's/.*release //')"
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
echo
echo "whole library:"
for m in UTCHMMA UTCHMMA.2CTA UTMALDG UTMALDG.2D.2CTA LDTM UTCBAR UTCBAR.2CTA.MULTICAST UTCATOMSWS SYNCS.EXCH SYNCS.ARRIVE HMMA FFMA; do
  printf "  %-24s %6d\n" "$m" "$(grep -c -- "$m" "$TMP")"
done
echo
echo "per tensor_scan_kernel instantiation (KP, L2, LIST, QRES, PAIR):"
awk '/Function : /{name=$3} /UTCHMMA/{u[name]++} /UTMALDG/{t[name]++} /LDTM/{l[name]++} /UTCBAR/{b[name]++} END{for(n in u) printf "  UTCHMMA=%-3d UTMALDG=%-3d LDTM=%-3d UTCBAR=%-3d %s\n", u[n], t[n], l[n], b[n], n}' "$TMP" | c++filt | sed 's/(CUtensorMap_st, CUtensorMap_st.*//' | sort -k5
rm -f "$TMP"
