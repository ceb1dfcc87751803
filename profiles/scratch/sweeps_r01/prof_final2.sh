#!/bin/bash
# end-of-round refresh: tests, smoke, both bench arms, frame API numbers, launch list of one linked-frame round trip
cd "$(dirname "$0")/../../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final2.log 2>&1 || exit 1
tail -1 gpurun_out/bench_final2.log | cut -c1-160
python bench.py --impl reference > gpurun_out/bench_ref_final2.log 2>&1
tail -1 gpurun_out/bench_ref_final2.log | cut -c1-160
python divortio-lz4_b200/tools/frame_bench.py log 64 > gpurun_out/frame_final2_log64.log 2>&1
python divortio-lz4_b200/tools/frame_bench.py mixed 1024 > gpurun_out/frame_final2_mixed1024.log 2>&1
python divortio-lz4_b200/tools/frame_bench.py log 1024 --only=0 --only=1 --only=2 > gpurun_out/frame_final2_log1024.log 2>&1
DLZ4_SEG_OVERLAP_MIN_MIB=100000 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_frame_final2.csv python divortio-lz4_b200/tools/frame_bench.py mixed 1024 --only=0 --once > gpurun_out/ncu_final2_3.log 2>&1
tail -2 gpurun_out/frame_final2_mixed1024.log | cut -c1-170
