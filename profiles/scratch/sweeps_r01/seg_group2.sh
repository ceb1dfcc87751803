#!/bin/bash
cd "$(dirname "$0")/../../.."
run() { echo "== $*"; env "$@" timeout 200 python divortio-lz4_b200/tools/frame_bench.py log $M --only=0 --only=1 2>&1 | grep block | cut -c1-170; }
M=1024
run DLZ4_X=default
run DLZ4_SEG_KIB=256 DLZ4_SEG_GROUP=2
run DLZ4_SEG_KIB=256 DLZ4_SEG_GROUP=3
run DLZ4_SEG_KIB=1024 DLZ4_SEG_GROUP=2
M=256
run DLZ4_X=default
run DLZ4_SEG_KIB=256 DLZ4_SEG_GROUP=2
M=192
run DLZ4_X=default
run DLZ4_SEG_GROUP=0
