#!/bin/bash
# A/B: decoder with / without the L1 prefetch of every real token's match source
L=divortio-lz4_b200/csrc
for lib in libdlz4_b200_nopf.so libdlz4_b200.so libdlz4_b200_nopf.so libdlz4_b200.so; do
  echo "== $lib"
  DLZ4_LIB=$PWD/$L/$lib timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed 2>&1 | tail -2
done
