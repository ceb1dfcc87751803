#!/bin/bash
# segment groups (DLZ4_SEG_GROUP) on linked 4 MiB-block frames
cd "$(dirname "$0")/../../.."
for m in 256 512 1024; do
  for g in 0 default 2 4 8; do
    if [ $g = default ]; then unset DLZ4_SEG_GROUP; else export DLZ4_SEG_GROUP=$g; fi
    echo "== $m MiB group=$g"
    timeout 200 python divortio-lz4_b200/tools/frame_bench.py log $m --only=0 2>&1 | grep block | cut -c1-170
  done
done
