#!/bin/bash
# End-of-round captures (one B200): tests, bench lines, ncu launch list of the bench, ncu --set full of the three hot kernels.
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo rc=$? >> gpurun_out/f_tests.log
timeout 900 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo bench rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
for k in k_parse_pw k_encode_blocks k_decompress_blocks; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/f_$k \
      python divortio-lz4_b200/tools/prof_one.py mixed 1024 > gpurun_out/f_ncu_$k.log 2>&1
done
timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 > gpurun_out/f_kbench.txt 2>&1
ls -la gpurun_out/
