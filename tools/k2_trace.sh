#!/bin/bash
# Builds lib/libb200flat_trace.so: the library with -DB2F_K2_TRACE (per-CTA phase stamps in K2; diagnostics only, the
# shipped library has none of it).  Used by tools/k2_trace.py through B200FLAT_LIB.
set -e
cd "$(dirname "$0")/../rag-faiss-embedding_b200"
mkdir -p lib/obj_trace
for f in csrc/*.cu; do
  o=lib/obj_trace/$(basename ${f%.cu}).o
  if [ ! -f $o ] || [ $f -nt $o ] || [ -n "$(find csrc ../include -newer $o \( -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -I ../include -DB2F_K2_TRACE -c $f -o $o &
  fi
done
wait
nvcc -shared -o lib/libb200flat_trace.so lib/obj_trace/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo lib/libb200flat_trace.so
