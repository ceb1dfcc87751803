#!/bin/bash
# decoder queue order: blocks that look like work first.  _nopf = the library of the commit before (no prefetch, one pass)
L=divortio-lz4_b200/csrc
for lib in libdlz4_b200_nopf.so libdlz4_b200.so libdlz4_b200.so; do
  echo "== $lib"
  DLZ4_LIB=$PWD/$L/$lib timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,zero,rand 2>&1 | tail -4
done
