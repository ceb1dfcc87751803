#!/bin/bash
# A/B: helper chains (k_compress_help) beside the k_parse_pw teams.  Usage: bash profiles/scratch/r02_help_sweep.sh > out.txt
for cfg in "0 0" "5 3072" "5 2048" "4 3072" "6 3072" "8 4096"; do
  set -- $cfg
  echo "== DLZ4_HELP=$1 DLZ4_HELP_GUARD=$2"
  DLZ4_HELP=$1 DLZ4_HELP_GUARD=$2 timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed 2>&1 | tail -3
done
