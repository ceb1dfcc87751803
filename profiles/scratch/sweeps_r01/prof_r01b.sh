mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_b.log 2>&1 || exit 1
tail -1 gpurun_out/plain_bench_b.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_compress_fresh16h -c 1 -s 3 -o gpurun_out/prof_b_compress -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_decompress_blocks -c 1 -s 3 -o gpurun_out/prof_b_decompress -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b2.log 2>&1
python divortio-lz4_b200/tools/frame_bench.py log 64 --no-cpu > gpurun_out/plain_frame.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:k_jd_scan -c 1 -o gpurun_out/prof_b_jdscan -f python divortio-lz4_b200/tools/frame_bench.py log 64 --no-cpu > gpurun_out/ncu_b3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_frame.csv python divortio-lz4_b200/tools/frame_bench.py log 64 --no-cpu > gpurun_out/ncu_b4.log 2>&1
ls -la gpurun_out | tail -12
