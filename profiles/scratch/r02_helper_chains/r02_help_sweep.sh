#!/bin/bash
# A/B: helper chains (k_compress_help) beside the k_parse_pw teams; lb800 = k_parse_pw capped at 72 registers per thread.
# Usage: bash profiles/scratch/r02_help_sweep.sh > out.txt
L=divortio-lz4_b200/csrc
for cfg in "$L/libdlz4_b200.so 0 0" "$L/libdlz4_b200.so 4 3072" "$L/libdlz4_b200.so 5 3072" "$L/libdlz4_b200_lb800.so 0 0" "$L/libdlz4_b200_lb800.so 5 3072" "$L/libdlz4_b200_lb800.so 8 3072" "$L/libdlz4_b200_lb800.so 8 2048"; do
  set -- $cfg
  echo "== $(basename $1) DLZ4_HELP=$2 DLZ4_HELP_GUARD=$3"
  DLZ4_LIB=$PWD/$1 DLZ4_HELP=$2 DLZ4_HELP_GUARD=$3 timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed 2>&1 | tail -3
done
