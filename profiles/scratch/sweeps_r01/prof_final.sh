mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2>&1 || exit 1
tail -1 gpurun_out/bench_final.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final0.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_compress_fresh16h -c 1 -s 3 -o gpurun_out/prof_final_compress -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_decompress_blocks -c 1 -s 3 -o gpurun_out/prof_final_decompress -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final2.log 2>&1
python divortio-lz4_b200/tools/frame_bench.py log 64 > gpurun_out/frame_final_log64.log 2>&1
python divortio-lz4_b200/tools/frame_bench.py mixed 1024 > gpurun_out/frame_final_mixed1024.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_frame_final.csv python divortio-lz4_b200/tools/frame_bench.py mixed 1024 --only=0 --once > gpurun_out/ncu_final3.log 2>&1
python divortio-lz4_b200/tools/config_bench.py 1024 262144 > gpurun_out/config_final.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1
tail -1 gpurun_out/bench_ref_final.log | cut -c1-200
ls gpurun_out | tail -5
